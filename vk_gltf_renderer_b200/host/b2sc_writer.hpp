// b2sc_writer.hpp -- writes the "B2SC" scene blob from the C-ABI structures of include/b200pt.h (header only, no other dependency).
//
// "scene-blob writer inside the reference loader".  A maintainer of the reference fills the same
// b200pt_scene_desc the shim of INTEGRATION.md section 2 hands to b200pt_set_scene -- GltfRenderNode[] as SceneVk uploads them
// (src/gltf_scene_vk.cpp:493-501), the per-primitive attribute arrays (:741-869), MaterialCache::getShadeMaterials() /
// getTextureInfos() (src/gltf_material_cache.hpp:76-77), the decoded RGBA8 images with their glTF sampler enums (:909-947,
// :1102-1154), getShaderLights() (:1354-1394) -- and calls writeB2sc: the file is what `b200pt_headless --scene x.b2sc` and
// SceneData::load read, so the reference's own loader, not this repo's re-parser, decides what the CUDA path renders.
// The layout is the one vk_gltf_renderer_b200/scene.py::Scene.save_blob documents; tests/test_host.py checks that a blob loaded by
// SceneData and written back by this header is byte-identical.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/b200pt.h"

namespace b200host {

struct BlobCamera  // what nvutils::CameraManipulator hands the renderer: look-at + lens (same fields as b200host::Camera)
{
  uint32_t orthographic = 0;
  float    eye[3] = {0, 0, 1}, center[3] = {0, 0, 0}, up[3] = {0, 1, 0};
  float    yfov = 0.785398f, znear = 0.1f, zfar = 1000.f, xmag = 1.f, ymag = 1.f;
};

inline void writeB2sc(const std::string& path, const b200pt_scene_desc& d, const BlobCamera& cam, const float* hdrRgb = nullptr, uint32_t hdrWidth = 0,
                      uint32_t hdrHeight = 0, const b200pt_micromap* micromaps = nullptr, uint32_t numMicromaps = 0,
                      const b200pt_primitive_omm* primOmms = nullptr, uint32_t numPrimOmms = 0)
{
  FILE* f = std::fopen(path.c_str(), "wb");
  if(!f)
    throw std::runtime_error("cannot write " + path);
  auto put = [&](const void* p, size_t n) {
    if(n && std::fwrite(p, 1, n, f) != n)
    {
      std::fclose(f);
      throw std::runtime_error("short write to " + path);
    }
  };
  auto u32 = [&](uint32_t v) { put(&v, 4); };
  put("B2SC", 4);
  u32(1);  // version
  u32(d.numRenderNodes), u32(d.numRenderPrimitives), u32(d.numMaterials), u32(d.numTextureInfos), u32(d.numTextures), u32(d.numLights);
  u32(cam.orthographic);
  const float c[14] = {cam.eye[0], cam.eye[1], cam.eye[2], cam.center[0], cam.center[1], cam.center[2], cam.up[0], cam.up[1], cam.up[2],
                       cam.yfov,   cam.znear,  cam.zfar,   cam.xmag,      cam.ymag};
  put(c, sizeof(c));
  for(uint32_t i = 0; i < d.numRenderNodes; i++)
  {
    const b200pt_render_node& n = d.renderNodes[i];
    put(n.objectToWorld, 64);
    put(n.worldToObject, 64);
    put(&n.materialID, 4);
    put(&n.renderPrimID, 4);
    u32(d.renderNodeVisible ? (d.renderNodeVisible[i] ? 1u : 0u) : 1u);
  }
  for(uint32_t i = 0; i < d.numRenderPrimitives; i++)
  {
    const b200pt_render_primitive& p = d.renderPrimitives[i];
    // attribute mask: bit 0 normals, 1 uv0, 2 uv1, 3 tangents, 4 colours -- and the arrays follow in that order
    const uint32_t mask = (p.normals ? 1u : 0u) | (p.texCoords[0] ? 2u : 0u) | (p.texCoords[1] ? 4u : 0u) | (p.tangents ? 8u : 0u) | (p.colors ? 16u : 0u);
    u32(p.vertexCount), u32(p.triangleCount), u32(mask);
    put(p.positions, (size_t)p.vertexCount * 12);
    put(p.indices, (size_t)p.triangleCount * 12);
    if(p.normals)
      put(p.normals, (size_t)p.vertexCount * 12);
    if(p.texCoords[0])
      put(p.texCoords[0], (size_t)p.vertexCount * 8);
    if(p.texCoords[1])
      put(p.texCoords[1], (size_t)p.vertexCount * 8);
    if(p.tangents)
      put(p.tangents, (size_t)p.vertexCount * 16);
    if(p.colors)
      put(p.colors, (size_t)p.vertexCount * 4);
  }
  put(d.materials, (size_t)d.numMaterials * sizeof(b200pt_shade_material));
  put(d.textureInfos, (size_t)d.numTextureInfos * sizeof(b200pt_texture_info));
  for(uint32_t i = 0; i < d.numTextures; i++)
  {
    const b200pt_texture& t = d.textures[i];
    const int32_t         hd[7] = {t.width, t.height, t.srgb, t.wrapS, t.wrapT, t.magFilter, t.minFilter};
    put(hd, sizeof(hd));
    put(t.rgba8, (size_t)t.width * (size_t)t.height * 4);
  }
  put(d.lights, (size_t)d.numLights * sizeof(b200pt_light));
  if(hdrRgb && hdrWidth && hdrHeight)
  {
    u32(hdrWidth), u32(hdrHeight);
    put(hdrRgb, (size_t)hdrWidth * hdrHeight * 12);
  }
  else
    u32(0), u32(0);
  if(numPrimOmms)  // optional trailing section: the EXT_mesh_opacity_micromap arrays (SceneOmm's input)
  {
    put("OMM1", 4);
    u32(numMicromaps), u32(numPrimOmms);
    for(uint32_t i = 0; i < numMicromaps; i++)
    {
      const b200pt_micromap& m = micromaps[i];
      put(&m.dataSize, 8);
      u32(m.numTriangles);
      put(m.data, (size_t)m.dataSize);
      put(m.triangles, (size_t)m.numTriangles * sizeof(b200pt_micromap_triangle));
    }
    for(uint32_t i = 0; i < numPrimOmms; i++)
    {
      const b200pt_primitive_omm& po = primOmms[i];
      u32(po.renderPrimID), u32(po.micromap), u32(po.baseTriangle), u32(po.indices ? po.numIndices : 0u);
      if(po.indices)
        put(po.indices, (size_t)po.numIndices * 4);
    }
  }
  std::fclose(f);
}

}  // namespace b200host
