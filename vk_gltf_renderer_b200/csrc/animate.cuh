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
