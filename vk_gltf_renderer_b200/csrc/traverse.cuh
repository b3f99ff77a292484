// traverse.cuh — software traversal of the compressed 8-wide BVH (see bvh.h for the layout).
//
// Replaces the hardware TLAS/BLAS traversal behind RayQuery / TraceRay in the reference
// (shaders/raytracer_interface.h.slang:69-122 Trace, :139-187 TraceShadow).  One query primitive
// serves both: "closest hit whose (t, globalTriangleId) is lexicographically greater than a lower
// bound".  The any-hit loops of the reference (stochastic alpha, coloured transmission) become
// front-to-back iterations of that query, which makes the result independent of tree layout.
#pragma once
#include "bvh.h"
#include "vec.cuh"

namespace pt {

struct TraceHit
{
  float    t;
  float    u, v;
  uint32_t slot;    // index into the triangle array, 0xFFFFFFFF = miss
  uint32_t gid;     // global (flatten-order) triangle id, tie-break key
  uint32_t w0;      // rnode | flags << 28
};

struct BvhView
{
  const float4* __restrict__ nodes;  // 5 per node
  const float4* __restrict__ tris;   // 3 per triangle
};

PT_D uint32_t extractByte(uint32_t x, int i) { return (x >> (i * 8)) & 0xffu; }
PT_D uint32_t signExtendS8x4(uint32_t x)
{
  // per byte: 0x80 -> 0xff, else 0x00  (vabsdiff4 trick replaced by plain bit math)
  return ((x >> 7) & 0x01010101u) * 0xffu;
}

#ifdef B200PT_COUNT_TRAVERSAL
#define PT_COUNT_NODE() (nodeCount++)
#define PT_COUNT_TRI() (triCount++)
#else
#define PT_COUNT_NODE()
#define PT_COUNT_TRI()
#endif

// MODE_CLOSEST: back-face culling per triangle flags (RAY_FLAG_CULL_BACK_FACING_TRIANGLES + instance
//               cull-disable);  MODE_SHADOW: no culling (RAY_FLAG_NONE).
// anyHitExit: stop at the first accepted triangle (valid only when every triangle is opaque).
template <bool CULL, bool ANY_EXIT>
PT_D TraceHit traverseNext(const BvhView bvh, float3 org, float3 dir, float tmin, float tmax, bool haveLo, float loT, uint32_t loId,
                           unsigned long long* nodeCounter = nullptr, unsigned long long* triCounter = nullptr)
{
#ifdef B200PT_COUNT_TRAVERSAL
  unsigned int nodeCount = 0, triCount = 0;
#endif
  TraceHit best;
  best.t = tmax;
  best.slot = 0xFFFFFFFFu;
  best.gid = 0xFFFFFFFFu;
  best.u = best.v = 0.f;
  best.w0 = 0;

  const float ooeps = 1e-20f;
  const float dx = fabsf(dir.x) > ooeps ? dir.x : copysignf(ooeps, dir.x);
  const float dy = fabsf(dir.y) > ooeps ? dir.y : copysignf(ooeps, dir.y);
  const float dz = fabsf(dir.z) > ooeps ? dir.z : copysignf(ooeps, dir.z);
  const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
  const uint32_t octInv = ((dir.x < 0.f ? 0u : 4u) | (dir.y < 0.f ? 0u : 2u) | (dir.z < 0.f ? 0u : 1u));
  const uint32_t octInv4 = octInv * 0x01010101u;
  const float    tLow = haveLo ? fmaxf(tmin, loT) : tmin;

  uint2 stack[32];
  int   sp = 0;
  uint2 cur = make_uint2(0u, 0x80000000u);

  while(true)
  {
    uint2 triGroup;
    if(cur.y & 0xff000000u)
    {
      const uint32_t hitsImask = cur.y;
      const int      childBit = 31 - __clz(hitsImask);
      cur.y &= ~(1u << childBit);
      if(cur.y & 0xff000000u)
      {
        if(sp < 32)
          stack[sp++] = cur;
      }
      const uint32_t slotIndex = (uint32_t)(childBit - 24) ^ (octInv4 & 0xffu);
      const uint32_t relative = __popc(hitsImask & ~(0xffffffffu << slotIndex));
      const uint32_t nodeIndex = cur.x + relative;
      PT_COUNT_NODE();

      const float4 n0 = __ldg(&bvh.nodes[nodeIndex * 5 + 0]);
      const float4 n1 = __ldg(&bvh.nodes[nodeIndex * 5 + 1]);
      const float4 n2 = __ldg(&bvh.nodes[nodeIndex * 5 + 2]);
      const float4 n3 = __ldg(&bvh.nodes[nodeIndex * 5 + 3]);
      const float4 n4 = __ldg(&bvh.nodes[nodeIndex * 5 + 4]);

      const uint32_t eImask = __float_as_uint(n0.w);
      const float    adx = __uint_as_float(extractByte(eImask, 0) << 23) * idx;
      const float    ady = __uint_as_float(extractByte(eImask, 1) << 23) * idy;
      const float    adz = __uint_as_float(extractByte(eImask, 2) << 23) * idz;
      const float    aox = (n0.x - org.x) * idx;
      const float    aoy = (n0.y - org.y) * idy;
      const float    aoz = (n0.z - org.z) * idz;

      cur.x = __float_as_uint(n1.x);
      triGroup.x = __float_as_uint(n1.y);

      uint32_t hitMask = 0;
#pragma unroll
      for(int i = 0; i < 2; i++)
      {
        const uint32_t meta4 = __float_as_uint(i == 0 ? n1.z : n1.w);
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = signExtendS8x4(isInner4 << 3);
        const uint32_t bitIndex4 = (meta4 ^ (octInv4 & innerMask4)) & 0x1f1f1f1fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;

        const uint32_t qlox = __float_as_uint(i == 0 ? n2.x : n2.y), qhix = __float_as_uint(i == 0 ? n2.z : n2.w);
        const uint32_t qloy = __float_as_uint(i == 0 ? n3.x : n3.y), qhiy = __float_as_uint(i == 0 ? n3.z : n3.w);
        const uint32_t qloz = __float_as_uint(i == 0 ? n4.x : n4.y), qhiz = __float_as_uint(i == 0 ? n4.z : n4.w);
        const uint32_t xmin = dir.x < 0.f ? qhix : qlox, xmax = dir.x < 0.f ? qlox : qhix;
        const uint32_t ymin = dir.y < 0.f ? qhiy : qloy, ymax = dir.y < 0.f ? qloy : qhiy;
        const uint32_t zmin = dir.z < 0.f ? qhiz : qloz, zmax = dir.z < 0.f ? qloz : qhiz;
#pragma unroll
        for(int j = 0; j < 4; j++)
        {
          const float tminx = fmaf((float)extractByte(xmin, j), adx, aox);
          const float tminy = fmaf((float)extractByte(ymin, j), ady, aoy);
          const float tminz = fmaf((float)extractByte(zmin, j), adz, aoz);
          const float tmaxx = fmaf((float)extractByte(xmax, j), adx, aox);
          const float tmaxy = fmaf((float)extractByte(ymax, j), ady, aoy);
          const float tmaxz = fmaf((float)extractByte(zmax, j), adz, aoz);
          const float tn = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, tLow));
          const float tf = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, best.t));
          // widen by a few ulp: keeps the box test conservative w.r.t. the triangle test
          if(tn <= tf * 1.000001f)
          {
            const uint32_t childBits = extractByte(childBits4, j);
            const uint32_t bitIndex = extractByte(bitIndex4, j);
            hitMask |= childBits << bitIndex;
          }
        }
      }
      cur.y = (hitMask & 0xff000000u) | (eImask >> 24);
      triGroup.y = hitMask & 0x00ffffffu;
    }
    else
    {
      triGroup = cur;
      cur = make_uint2(0u, 0u);
    }

    while(triGroup.y != 0)
    {
      const int triBit = 31 - __clz(triGroup.y);
      triGroup.y &= ~(1u << triBit);
      const uint32_t slot = triGroup.x + (uint32_t)triBit;
      PT_COUNT_TRI();
      const float4 a = __ldg(&bvh.tris[slot * 3 + 0]);
      const float4 b = __ldg(&bvh.tris[slot * 3 + 1]);
      const float4 c = __ldg(&bvh.tris[slot * 3 + 2]);
      const float3 v0 = f3(a.x, a.y, a.z), e1 = f3(b.x, b.y, b.z), e2 = f3(c.x, c.y, c.z);
      // Moeller-Trumbore, explicit fma chain (bit-identical to oracle/pt_oracle.cpp intersectTri)
      const float3 pvec = crossFma(dir, e2);
      const float  det = dotFma(e1, pvec);
      if(det == 0.0f)
        continue;
      const float  inv = 1.0f / det;
      const float3 tvec = org - v0;
      const float  u = dotFma(tvec, pvec) * inv;
      if(u < 0.0f || u > 1.0f)
        continue;
      const float3 qvec = crossFma(tvec, e1);
      const float  v = dotFma(dir, qvec) * inv;
      if(v < 0.0f || u + v > 1.0f)
        continue;
      const float    t = dotFma(e2, qvec) * inv;
      const uint32_t w0 = __float_as_uint(a.w);
      const uint32_t flags = w0 >> 28;
      if(CULL && !(flags & TRI_NOCULL))
      {
        const bool front = (flags & TRI_FLIPPED) ? (det < 0.0f) : (det > 0.0f);
        if(!front)
          continue;
      }
      if(!(t > tmin && t < tmax))
        continue;
      const uint32_t gid = __float_as_uint(c.w);
      if(haveLo && !(t > loT || (t == loT && gid > loId)))
        continue;
      if(t < best.t || (t == best.t && gid < best.gid))
      {
        best.t = t;
        best.u = u;
        best.v = v;
        best.slot = slot;
        best.gid = gid;
        best.w0 = w0;
        if(ANY_EXIT)
        {
          sp = 0;
          cur.y = 0;
          break;
        }
      }
    }

    if((cur.y & 0xff000000u) == 0)
    {
      if(sp == 0)
        break;
      cur = stack[--sp];
    }
  }
  if(best.slot != 0xFFFFFFFFu && ((best.w0 >> 28) & TRI_FLIPPED))
  {
    const float tmp = best.u;
    best.u = best.v;
    best.v = tmp;
  }
#ifdef B200PT_COUNT_TRAVERSAL
  if(nodeCounter)
    atomicAdd(nodeCounter, (unsigned long long)nodeCount);
  if(triCounter)
    atomicAdd(triCounter, (unsigned long long)triCount);
#endif
  return best;
}

}  // namespace pt
