// bvh.h — host-side builder for the 8-wide compressed BVH (CWBVH, Ylitie/Karras/Laine 2017) the
// sm_100a traversal kernels walk.
//
// Replaces what the reference gets from the driver: one BLAS per render primitive + one TLAS
// instance per render node (src/gltf_scene_rtx.cpp:140-170, 299-388).  Scenes are static in every
// BASELINE config, so instances are flattened: every visible render node contributes its
// triangles in world space to ONE tree (no per-instance ray transform on the hot path; 180 GB of
// HBM makes the duplication of instanced meshes affordable).  Instance semantics are kept per
// triangle: rnode id (-> InstanceIndex), primitive id (-> PrimitiveIndex), opaque / cull-disable
// flags from getInstanceFlag (src/gltf_scene_rtx.cpp:271-295).
#pragma once
#include <stdint.h>

#include <vector>

namespace pt {

enum : uint32_t
{
  TRI_OPAQUE = 1u,   // VK_GEOMETRY_INSTANCE_FORCE_OPAQUE: no any-hit work
  TRI_NOCULL = 2u,   // VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE
  TRI_FLIPPED = 4u,  // mirrored instance: world-space winding was swapped, (u,v) swap back
};

struct FlatTri
{
  float    v0[3], e1[3], e2[3];
  uint32_t rnode;
  uint32_t prim;
  uint32_t flags;
};

// Node = 80 B = 5 x 16 B:
//   n0 = (p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24)
//   n1 = (childBase | axisMap << 26, triBase, meta[0..3], meta[4..7])   axisMap: which axis each slot bit follows (bvh.cpp)
//   n2 = (qlo_x[0..3], qlo_x[4..7], qhi_x[0..3], qhi_x[4..7]);  n3 = y;  n4 = z
// Triangle = 48 B = 3 x 16 B: (v0.xyz, rnode | flags<<28) (e1.xyz, prim) (e2.xyz, globalId)
struct WideBvh
{
  std::vector<float>    nodes;  // 20 floats per node
  std::vector<float>    tris;   // 12 floats per triangle, leaf order
  std::vector<uint32_t> triMeta;  // 2 per triangle: rnode | flags<<28, prim  (shade-stage lookup)
  uint32_t              numNodes = 0, numTris = 0;
  float                 boundsLo[3], boundsHi[3];
  uint32_t              maxDepth = 0;
};

// gids[i] = global (flatten-order) id of tris[i], the tie-break key stored in the triangle record;
// triBaseOffset is added to every node's triangle base (several trees may share one triangle array).
void buildWideBvh(const std::vector<FlatTri>& tris, const std::vector<uint32_t>& gids, uint32_t triBaseOffset, WideBvh& out);

}  // namespace pt
