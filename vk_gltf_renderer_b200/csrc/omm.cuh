// omm.cuh — opacity micromaps in the software walk (device + host code).
//
// Reference: src/gltf_scene_omm.{hpp,cpp} uploads the asset's EXT_mesh_opacity_micromap arrays as VK_EXT_opacity_micromap build
// input, SceneRtx attaches the micromap to the BLAS geometry, and from then on the RT cores resolve a hit on an alpha-tested
// triangle from the micro-triangle's state: OPAQUE commits without an any-hit invocation, TRANSPARENT is culled, only the UNKNOWN
// states reach the shader's getOpacity + rand() (docs/RENDERING_ARCHITECTURE.md:65-78, raytracer_interface.h.slang:93-100).
// There are no RT cores here, so the traversal kernels do that lookup themselves at the moment a non-opaque triangle is hit:
//
//   ommRef[slot]  one word per triangle slot of a tree (b200pt_set_scene resolves render primitive -> micromap -> triangle record):
//                   bits 28-31  subdivision level 0..12, or 15 = no lookup, the state is in bits 0-1 (special indices, no micromap)
//                   bit  27     format: 1 = 4-state (2 bits per micro-triangle), 0 = 2-state (1 bit)
//                   bits 0-26   byte offset of the triangle's states in ommData
//   ommData       the micromaps' `data` arrays back to back, the asset's own bytes (micro-triangles in the Vulkan "bird curve" order)
//
// The micro-triangle index of a barycentric position is the VK_EXT_opacity_micromap specification's bary2index(), restated here.
// PARITY UNPINNED: no asset carrying the extension (and no Vulkan driver) is reachable from this build, so the restatement is held
// only by its properties -- for every level a bijection of the 4^level micro-triangles, hierarchical (index >> 2 is the parent's
// index one level up), level 1 = {corner w, centre, corner u, corner v} (tests/test_omm.py) -- and by the baker, the oracle and
// these kernels all using the same order.
#pragma once
#include "vec.cuh"
#include <cstdint>

namespace pt {

constexpr uint32_t kOmmNoLookup = 15u << 28;  // | state
constexpr int      OMM_TRANSPARENT = 0, OMM_OPAQUE = 1, OMM_UNKNOWN = 2;

PT_HD uint32_t ommSpreadBits(uint32_t x)
{
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

// micro-triangle index of barycentrics (u, v) = weights of the triangle's 2nd and 3rd vertex, at a subdivision level
PT_HD uint32_t ommBary2Index(float u, float v, uint32_t level)
{
  u = fminf(fmaxf(u, 0.0f), 1.0f);
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  const uint32_t n = 1u << level;
  const float    fu = u * (float)n, fv = v * (float)n;
  uint32_t       iu = (uint32_t)fu, iv = (uint32_t)fv;
  const float    uf = fu - (float)iu, vf = fv - (float)iv;
  iu = iu >= n ? n - 1u : iu;
  iv = iv >= n ? n - 1u : iv;
  const uint32_t iuv = iu + iv;
  if(iuv >= n)
    iu -= iuv - n + 1u;
  uint32_t iw = ~(iu + iv);
  if(uf + vf >= 1.0f && iuv < n - 1u)
    --iw;
  const uint32_t b0 = ~(iu ^ iw) & (n - 1u);
  const uint32_t t = (iu ^ iv) & b0;
  uint32_t       f = t;
  f ^= f >> 1;
  f ^= f >> 2;
  f ^= f >> 4;
  f ^= f >> 8;
  const uint32_t b1 = ((f ^ iu) & ~b0) | t;
  return ommSpreadBits(b0) | (ommSpreadBits(b1) << 1);
}

// state of the micro-triangle under (u, v): OMM_TRANSPARENT / OMM_OPAQUE / OMM_UNKNOWN
template <typename LoadByte>
PT_HD int ommStateOf(uint32_t ref, float u, float v, LoadByte loadByte)
{
  const uint32_t level = ref >> 28;
  if(level == 15u)
    return (int)(ref & 3u);
  const uint32_t idx = ommBary2Index(u, v, level);
  const uint32_t off = ref & 0x07ffffffu;
  if(ref & (1u << 27))
  {
    const uint32_t s = (loadByte(off + (idx >> 2)) >> ((idx & 3u) * 2u)) & 3u;
    return s < 2u ? (int)s : OMM_UNKNOWN;
  }
  return (int)((loadByte(off + (idx >> 3)) >> (idx & 7u)) & 1u);
}

}  // namespace pt
