// animate.cuh — the animation feed: morph-target blending and skeletal skinning of a render primitive's vertex arrays
// on the device, followed by the re-gather of the per-triangle shade records (the caller then refits the trees, refit.cuh).
//
// Reference: shaders/morph.comp.slang:29-70, shaders/skinning.comp.slang:27-70 (push constants: shaders/animation_io.h.slang:29-59)
// dispatched by SceneAnimationVk::cmdUpdateAnimation (src/gltf_scene_animation_vk.cpp:396-592): morph first, skin second; a
// primitive that is both morphed and skinned is skinned from the morphed vertex buffers (:545-556); outputs go straight into
// the primitive's position / normal / tangent vertex buffers, which the BLAS update then reads.
//
// Matrix convention: the host hands glm column-major bytes; the shaders' mul(v, M) is M_glm * v.  The reference
// compiles these products through SPIR-V (operation order and contraction are the driver's); pinned here, in the device code
// and in oracle/animation.py alike: every product and sum is a separate fp32 operation (no FMA: the TU is built with
// -fmad=false), a row is ((m0 * x + m1 * y) + m2 * z) [+ m3], normalize(v) = v / sqrt((x*x + y*y) + z*z).
#pragma once
#include "device_scene.cuh"

namespace pt {

struct MorphTaskDev
{
  const float *basePos, *baseNrm, *baseTan;  // vertexCount x 3 / x 3 / x 4 (nullable: normals, tangents)
  const float *dPos, *dNrm, *dTan;           // numTargets x vertexCount x 3 (nullable: normals, tangents)
  const float* weights;                      // numTargets
  float *      outPos, *outNrm, *outTan;     // the primitive's vertex arrays (nullable: normals, tangents)
  uint32_t     vertexCount, numTargets;
};

struct SkinTaskDev
{
  const float *basePos, *baseNrm, *baseTan;  // static base arrays, or the primitive's own arrays when it was morphed this frame
  const float* weights;                      // vertexCount x 4
  const int*   joints;                       // vertexCount x 4
  const float* jointMatrices;                // numJoints x 16, glm mat4 bytes
  const float* normalMatrices;               // numJoints x 9, glm mat3 bytes
  float *      outPos, *outNrm, *outTan;
  uint32_t     vertexCount, numJoints;
};

PT_D float3 normalizeDiv(float3 v)
{
  const float l = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
  return f3(v.x / l, v.y / l, v.z / l);
}

// morph.comp.slang:29-70
PT_D void morphVertex(const MorphTaskDev& T, uint32_t v)
{
  const bool hasNormals = T.baseNrm != nullptr, hasTangents = T.baseTan != nullptr;
  float3     pos = f3(T.basePos[v * 3], T.basePos[v * 3 + 1], T.basePos[v * 3 + 2]);
  float3     nrm = hasNormals ? f3(T.baseNrm[v * 3], T.baseNrm[v * 3 + 1], T.baseNrm[v * 3 + 2]) : f3(0.f, 0.f, 0.f);
  float3     tan = hasTangents ? f3(T.baseTan[v * 4], T.baseTan[v * 4 + 1], T.baseTan[v * 4 + 2]) : f3(0.f, 0.f, 0.f);
  const float tanW = hasTangents ? T.baseTan[v * 4 + 3] : 0.f;
  for(uint32_t t = 0; t < T.numTargets; t++)
  {
    const float w = T.weights[t];
    if(w == 0.0f)
      continue;
    const size_t o = ((size_t)t * T.vertexCount + v) * 3;
    pos = f3(pos.x + w * T.dPos[o], pos.y + w * T.dPos[o + 1], pos.z + w * T.dPos[o + 2]);
    if(hasNormals && T.dNrm != nullptr)
      nrm = f3(nrm.x + w * T.dNrm[o], nrm.y + w * T.dNrm[o + 1], nrm.z + w * T.dNrm[o + 2]);
    if(hasTangents && T.dTan != nullptr)
      tan = f3(tan.x + w * T.dTan[o], tan.y + w * T.dTan[o + 1], tan.z + w * T.dTan[o + 2]);
  }
  T.outPos[v * 3] = pos.x, T.outPos[v * 3 + 1] = pos.y, T.outPos[v * 3 + 2] = pos.z;
  if(hasNormals && T.outNrm != nullptr)
  {
    const float3 n = normalizeDiv(nrm);
    T.outNrm[v * 3] = n.x, T.outNrm[v * 3 + 1] = n.y, T.outNrm[v * 3 + 2] = n.z;
  }
  if(hasTangents && T.outTan != nullptr)
  {
    const float3 t = normalizeDiv(tan);
    T.outTan[v * 4] = t.x, T.outTan[v * 4 + 1] = t.y, T.outTan[v * 4 + 2] = t.z, T.outTan[v * 4 + 3] = tanW;
  }
}

// skinning.comp.slang:27-70
PT_D void skinVertex(const SkinTaskDev& T, uint32_t v)
{
  const bool hasNormals = T.baseNrm != nullptr, hasTangents = T.baseTan != nullptr;
  // read everything of this vertex first: base and output may be the same arrays (morph -> skin composition)
  const float3 p = f3(T.basePos[v * 3], T.basePos[v * 3 + 1], T.basePos[v * 3 + 2]);
  const float3 n = hasNormals ? f3(T.baseNrm[v * 3], T.baseNrm[v * 3 + 1], T.baseNrm[v * 3 + 2]) : f3(0.f, 0.f, 0.f);
  const float3 t = hasTangents ? f3(T.baseTan[v * 4], T.baseTan[v * 4 + 1], T.baseTan[v * 4 + 2]) : f3(0.f, 0.f, 0.f);
  const float  tanW = hasTangents ? T.baseTan[v * 4 + 3] : 0.f;
  float3       sp = f3(0.f, 0.f, 0.f), sn = f3(0.f, 0.f, 0.f), st = f3(0.f, 0.f, 0.f);
  for(int i = 0; i < 4; i++)
  {
    const float jw = T.weights[v * 4 + i];
    const int   ji = T.joints[v * 4 + i];
    if(jw > 0.0f && ji >= 0 && (uint32_t)ji < T.numJoints)
    {
      const float* M = T.jointMatrices + (size_t)ji * 16;
      const float3 q = xfPoint(M, p);
      sp = f3(sp.x + jw * q.x, sp.y + jw * q.y, sp.z + jw * q.z);
      if(hasNormals)
      {
        const float* N = T.normalMatrices + (size_t)ji * 9;  // glm mat3: column c at N[3c..3c+2]
        const float3 r = f3((N[0] * n.x + N[3] * n.y) + N[6] * n.z, (N[1] * n.x + N[4] * n.y) + N[7] * n.z, (N[2] * n.x + N[5] * n.y) + N[8] * n.z);
        sn = f3(sn.x + jw * r.x, sn.y + jw * r.y, sn.z + jw * r.z);
      }
      if(hasTangents)
      {
        const float3 r = xfVector(M, t);
        st = f3(st.x + jw * r.x, st.y + jw * r.y, st.z + jw * r.z);
      }
    }
  }
  T.outPos[v * 3] = sp.x, T.outPos[v * 3 + 1] = sp.y, T.outPos[v * 3 + 2] = sp.z;
  if(hasNormals && T.outNrm != nullptr)
  {
    const float3 r = normalizeDiv(sn);
    T.outNrm[v * 3] = r.x, T.outNrm[v * 3 + 1] = r.y, T.outNrm[v * 3 + 2] = r.z;
  }
  if(hasTangents && T.outTan != nullptr)
  {
    const float3 r = normalizeDiv(st);
    T.outTan[v * 4] = r.x, T.outTan[v * 4 + 1] = r.y, T.outTan[v * 4 + 2] = r.z, T.outTan[v * 4 + 3] = tanW;
  }
}

// ---- rigid part on the device: world matrices down the node hierarchy, then the render nodes -------------------------------
// shaders/world_matrix_propagate.comp.slang:27-42 (one dispatch per topological level: world[node] = world[parent] * local[node];
// the shader's mul(local, parent) on glm bytes is the glm product parent * local) and
// shaders/update_render_instances.comp.slang:42-66 (objectToWorld = mul(world[nodeID], instLocal[i]) -- on glm bytes that is the
// glm product instLocal * world, followed literally although the CPU path of the reference multiplies the other way round,
// src/gltf_scene.cpp:2419; the two agree whenever instLocal is the identity, i.e. without EXT_mesh_gpu_instancing --
// worldToObject = inverse(objectToWorld), materialID / renderPrimID from the mapping).  Matrices are glm column-major: m[4 c + r].
// Product: C[c][r] = ((A[0][r] B[c][0] + A[1][r] B[c][1]) + A[2][r] B[c][2]) + A[3][r] B[c][3] (glm's operator*); inverse: glm's
// cofactor expansion (compute_inverse<4,4>), one reciprocal of the determinant times the adjugate.  The reference's shader calls the
// GLSL.std.450 MatrixInverse, whose arithmetic is the driver's; pinned here and in oracle/animation.py alike.
struct RenderNodeMapping  // RenderNodeGpuMapping, shaders/world_matrix_io.h.slang:41-47
{
  int nodeID, pad0, materialID, renderPrimID;
};

PT_HD void mat4Mul(const float* A, const float* B, float* C)  // C = A * B, may not alias
{
  for(int c = 0; c < 4; c++)
    for(int r = 0; r < 4; r++)
      C[4 * c + r] = ((A[r] * B[4 * c] + A[4 + r] * B[4 * c + 1]) + A[8 + r] * B[4 * c + 2]) + A[12 + r] * B[4 * c + 3];
}

PT_HD void mat4Inverse(const float* m, float* o)
{
#define M(c, r) m[4 * (c) + (r)]
  const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
  const float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  const float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
  const float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3_[4] = {c12, c12, c14, c15}, f4_[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  float       inv[16];
  for(int k = 0; k < 4; k++)
  {
    const float sa = (k & 1) ? -1.0f : 1.0f, sb = -sa;
    inv[0 + k] = ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]) * sa;
    inv[4 + k] = ((v0[k] * f0[k] - v2[k] * f3_[k]) + v3[k] * f4_[k]) * sb;
    inv[8 + k] = ((v0[k] * f1[k] - v1[k] * f3_[k]) + v3[k] * f5[k]) * sa;
    inv[12 + k] = ((v0[k] * f2[k] - v1[k] * f4_[k]) + v2[k] * f5[k]) * sb;
  }
  const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
  const float oneOverDet = 1.0f / ((d0 + d1) + (d2 + d3));
  for(int k = 0; k < 16; k++)
    o[k] = inv[k] * oneOverDet;
#undef M
}

// one node of a level (world_matrix_propagate.comp.slang)
PT_D void propagateNode(const float* local, float* world, const int* parents, const int* topoOrder, uint32_t levelOffset, uint32_t ti)
{
  const int node = topoOrder[levelOffset + ti], parent = parents[node];
  float     out[16];
  if(parent < 0)
    for(int k = 0; k < 16; k++)
      out[k] = local[(size_t)node * 16 + k];  // identity * local: exact
  else
    mat4Mul(world + (size_t)parent * 16, local + (size_t)node * 16, out);
  for(int k = 0; k < 16; k++)
    world[(size_t)node * 16 + k] = out[k];
}

// one render node (update_render_instances.comp.slang; the TlasInstance row it also writes is the refit's job here)
PT_D void updateRenderNode(const float* world, const RenderNodeMapping* mappings, const float* instLocal, b200pt_render_node* out, uint32_t i)
{
  const RenderNodeMapping map = mappings[i];
  float                   w[16];
  if(instLocal)
    mat4Mul(instLocal + (size_t)i * 16, world + (size_t)map.nodeID * 16, w);
  else
    for(int k = 0; k < 16; k++)
      w[k] = world[(size_t)map.nodeID * 16 + k];
  b200pt_render_node rn;
  for(int k = 0; k < 16; k++)
    rn.objectToWorld[k] = w[k];
  mat4Inverse(w, rn.worldToObject);
  rn.materialID = map.materialID;
  rn.renderPrimID = map.renderPrimID;
  out[i] = rn;
}

// the position / normal / tangent part of one triangle's ShadeRec from the primitive's (updated) vertex arrays; texture
// coordinates, colours and the flags word keep what b200pt_set_scene gathered (layout: device_scene.cuh ShadeRec)
PT_D void regatherShadeRec(ShadeRec* rec, const DevPrim& P, uint32_t tri)
{
  float* f = reinterpret_cast<float*>(rec);
  for(int c = 0; c < 3; c++)
  {
    const uint32_t vi = P.idx[tri * 3 + c];
    f[c * 4] = P.pos[vi * 3], f[c * 4 + 1] = P.pos[vi * 3 + 1], f[c * 4 + 2] = P.pos[vi * 3 + 2];
    if(P.nrm)
      f[12 + c * 4] = P.nrm[vi * 3], f[13 + c * 4] = P.nrm[vi * 3 + 1], f[14 + c * 4] = P.nrm[vi * 3 + 2];
    if(P.tan)
      f[36 + c * 4] = P.tan[vi * 4], f[37 + c * 4] = P.tan[vi * 4 + 1], f[38 + c * 4] = P.tan[vi * 4 + 2], f[39 + c * 4] = P.tan[vi * 4 + 3];
  }
}

}  // namespace pt
