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

// one non-opaque candidate kept by a collecting traversal (see TravState::collectN)
struct Cand
{
  float    t, u, v;
  uint32_t slot, gid;
};
#ifndef B200PT_KCAND
#define B200PT_KCAND 4
#endif
constexpr int kCand = B200PT_KCAND;

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

// byte j of x as the float 32768 + b: PRMT puts the byte into bits 8..15 of 0x47000000 (= 32768.0f, ulp 2^-8)
PT_D float biasedByte(uint32_t x, int j) { return __uint_as_float(__byte_perm(x, 0x47000000u, 0x7604u | ((uint32_t)j << 4))); }

// One traversal in flight, resumable one node-step at a time (the persistent kernels interleave the
// steps of 32 independent rays per warp and re-fill finished lanes from a global work counter).
//   cull    : back-face culling per triangle flags (RAY_FLAG_CULL_BACK_FACING_TRIANGLES + instance cull-disable)
//   anyExit : stop at the first accepted triangle (occlusion query against the opaque tree)
//   lo      : only hits lexicographically after (loT, loId) in (t, global id) order count
struct TravState
{
  const float4* nodes;
  const float4* tris;
  float3        org, dir;
  float         idx, idy, idz;
  float         tmin, tmax, tLow, loT;
  uint32_t      loId, octInv4;
  bool          haveLo, cull, anyExit;
  TraceHit      best;
  uint2         cur;   // current node group (x = child base, y = hit bits << 24 | imask)
  uint2         tri;   // pending triangle group (x = triangle base, y = hit bits)
  int           sp;    // entries on the caller-provided stack (kept OUT of this struct so the rest stays in registers)
  // Collecting mode (any-hit candidates): instead of keeping only the nearest hit, the traversal keeps the kCand
  // nearest hits after the lower bound, sorted by (t, id), in caller-provided scratch.  The kernels then run the
  // stochastic alpha / transmission tests over them front to back -- the same sequence the restart-per-candidate
  // formulation produces, with one tree walk per kCand candidates instead of one per candidate.  `best` holds the
  // kCand-th candidate once the list is full, so node culling and the hit predicate need no extra code.
  int           collectN;  // -1: nearest-hit mode, else number of candidates collected so far
#ifdef B200PT_COUNT_TRAVERSAL
  unsigned int nodeCount, triCount;
#endif

  PT_D void init(const BvhView bvh, float3 o, float3 d, float tmin_, float tmax_, bool cull_, bool anyExit_, bool haveLo_, float loT_, uint32_t loId_,
                 bool collect_ = false)
  {
    collectN = collect_ ? 0 : -1;
    nodes = bvh.nodes;
    tris = bvh.tris;
    org = o;
    dir = d;
    const float ooeps = 1e-20f;
    const float dx = fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x);
    const float dy = fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y);
    const float dz = fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z);
    idx = 1.0f / dx;
    idy = 1.0f / dy;
    idz = 1.0f / dz;
    octInv4 = ((d.x < 0.f ? 0u : 4u) | (d.y < 0.f ? 0u : 2u) | (d.z < 0.f ? 0u : 1u)) * 0x01010101u;
    tmin = tmin_;
    tmax = tmax_;
    haveLo = haveLo_;
    loT = loT_;
    loId = loId_;
    tLow = haveLo_ ? fmaxf(tmin_, loT_) : tmin_;
    cull = cull_;
    anyExit = anyExit_;
    best.t = tmax_;
    best.slot = 0xFFFFFFFFu;
    best.gid = 0xFFFFFFFFu;
    best.u = best.v = 0.f;
    best.w0 = 0;
    cur = make_uint2(0u, 0x80000000u);
    tri = make_uint2(0u, 0u);
    sp = 0;
#ifdef B200PT_COUNT_TRAVERSAL
    nodeCount = triCount = 0;
#endif
  }

  static constexpr int kStackSize = 28;

  // One traversal step: (1) lanes without pending triangles open their next node, (2) lanes with pending
  // triangles test ONE triangle each — but only when enough lanes of the warp have one (vote); otherwise the
  // group is postponed onto the stack and node traversal continues (Ylitie et al. 2017, section 4.3).
  // Returns true when the traversal is complete.  `stack` is per-thread scratch of kStackSize entries.
  PT_D bool step(uint2* __restrict__ stack, int postponeShift = 2, Cand* __restrict__ cand = nullptr)
  {
    // single exit: an early return inside the divergent regions would move their reconvergence point out of the
    // caller's loop and the lanes of a warp would drift apart (measured: 8 of 32 lanes active)
    bool done = false;
    if(tri.y == 0)
    {
      if((cur.y & 0xff000000u) == 0)
      {
        if(sp == 0)
          done = true;
        else
          cur = stack[--sp];
      }
      if(done)
      {
      }
      else if(cur.y & 0xff000000u)
      {
        const uint32_t hitsImask = cur.y;
        const int      childBit = 31 - __clz(hitsImask);
        cur.y &= ~(1u << childBit);
        if(cur.y & 0xff000000u)
        {
          if(sp < kStackSize)
            stack[sp++] = cur;
        }
        const uint32_t slotIndex = (uint32_t)(childBit - 24) ^ (octInv4 & 0xffu);
        const uint32_t relative = __popc(hitsImask & ~(0xffffffffu << slotIndex));
        const uint32_t nodeIndex = cur.x + relative;
#ifdef B200PT_COUNT_TRAVERSAL
        nodeCount++;
#endif
        const float4 n0 = __ldg(&nodes[nodeIndex * 5 + 0]);
        const float4 n1 = __ldg(&nodes[nodeIndex * 5 + 1]);
        const float4 n2 = __ldg(&nodes[nodeIndex * 5 + 2]);
        const float4 n3 = __ldg(&nodes[nodeIndex * 5 + 3]);
        const float4 n4 = __ldg(&nodes[nodeIndex * 5 + 4]);

        const uint32_t eImask = __float_as_uint(n0.w);
        const float    adx = __uint_as_float(extractByte(eImask, 0) << 23) * idx;
        const float    ady = __uint_as_float(extractByte(eImask, 1) << 23) * idy;
        const float    adz = __uint_as_float(extractByte(eImask, 2) << 23) * idz;
        // Quantised plane bytes become floats WITHOUT the conversion unit (48 I2F.U8 per node saturated the XU
        // pipe, ncu: 65 % active, the busiest pipe): one PRMT drops byte b into the mantissa of 2^15, giving
        // 32768 + b exactly, and the bias moves into the addend: t = (32768 + b) * ad + (ao - 32768 * ad).
        // The addend's rounding error is at most |ad| / 512 (1/512 of a quantisation cell); the entry side is
        // pulled back and the exit side pushed out by |ad| / 256, so the test stays conservative.
        const float aox = fmaf(-32768.0f, adx, (n0.x - org.x) * idx);
        const float aoy = fmaf(-32768.0f, ady, (n0.y - org.y) * idy);
        const float aoz = fmaf(-32768.0f, adz, (n0.z - org.z) * idz);
        const float aoxN = fmaf(-0.00390625f, fabsf(adx), aox), aoxF = fmaf(0.00390625f, fabsf(adx), aox);
        const float aoyN = fmaf(-0.00390625f, fabsf(ady), aoy), aoyF = fmaf(0.00390625f, fabsf(ady), aoy);
        const float aozN = fmaf(-0.00390625f, fabsf(adz), aoz), aozF = fmaf(0.00390625f, fabsf(adz), aoz);

        cur.x = __float_as_uint(n1.x);
        tri.x = __float_as_uint(n1.y);

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
            const float tminx = fmaf(biasedByte(xmin, j), adx, aoxN);
            const float tminy = fmaf(biasedByte(ymin, j), ady, aoyN);
#ifdef B200PT_CVT_Z_I2F
            const float tminz = fmaf((float)extractByte(zmin, j), adz, fmaf(32768.0f, adz, aozN));
#else
            const float tminz = fmaf(biasedByte(zmin, j), adz, aozN);
#endif
            const float tmaxx = fmaf(biasedByte(xmax, j), adx, aoxF);
            const float tmaxy = fmaf(biasedByte(ymax, j), ady, aoyF);
#ifdef B200PT_CVT_Z_I2F
            const float tmaxz = fmaf((float)extractByte(zmax, j), adz, fmaf(32768.0f, adz, aozF));
#else
            const float tmaxz = fmaf(biasedByte(zmax, j), adz, aozF);
#endif
            const float tn = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, tLow));
            const float tf = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, best.t));
            // widen by a few ulp: keeps the box test conservative w.r.t. the triangle test
            const bool     in = tn <= tf * 1.000001f;
            const uint32_t childBits = in ? extractByte(childBits4, j) : 0u;
            hitMask |= childBits << extractByte(bitIndex4, j);
          }
        }
        cur.y = (hitMask & 0xff000000u) | (eImask >> 24);
        tri.y = hitMask & 0x00ffffffu;
      }
      else
      {
        // a postponed triangle group came off the stack
        tri = cur;
        cur = make_uint2(0u, 0u);
      }
    }

    // ---- triangle phase (one triangle per lane per step) ------------------------------------------------
    const unsigned conv = __activemask();
    const unsigned haveTri = __ballot_sync(conv, !done && tri.y != 0);
    if(!done && tri.y != 0)
    {
      // postpone when few lanes would take part and this lane's current node group still has children to open
      // (the group goes under the next node; `cur` only ever holds node groups)
      if((__popc(haveTri) << postponeShift) < __popc(conv) && (cur.y & 0xff000000u) != 0 && sp < kStackSize)
      {
        stack[sp++] = tri;
        tri.y = 0;
      }
      else
      {
        const int triBit = 31 - __clz(tri.y);
        tri.y &= ~(1u << triBit);
        const uint32_t slot = tri.x + (uint32_t)triBit;
#ifdef B200PT_COUNT_TRAVERSAL
        triCount++;
#endif
        const float4 a = __ldg(&tris[slot * 3 + 0]);
        const float4 b = __ldg(&tris[slot * 3 + 1]);
        const float4 c = __ldg(&tris[slot * 3 + 2]);
        const float3 v0 = f3(a.x, a.y, a.z), e1 = f3(b.x, b.y, b.z), e2 = f3(c.x, c.y, c.z);
        // Moeller-Trumbore, explicit fma chain (bit-identical to oracle/pt_oracle.cpp intersectTri); branch-free
        // up to the final update so the lanes stay converged
        const float3   pvec = crossFma(dir, e2);
        const float    det = dotFma(e1, pvec);
        const float    inv = 1.0f / det;
        const float3   tvec = org - v0;
        const float    u = dotFma(tvec, pvec) * inv;
        const float3   qvec = crossFma(tvec, e1);
        const float    v = dotFma(dir, qvec) * inv;
        const float    t = dotFma(e2, qvec) * inv;
        const uint32_t w0 = __float_as_uint(a.w);
        const uint32_t flags = w0 >> 28;
        const uint32_t gid = __float_as_uint(c.w);
        bool           hit = (det != 0.0f) & (u >= 0.0f) & (u <= 1.0f) & (v >= 0.0f) & (u + v <= 1.0f) & (t > tmin) & (t < tmax);
        const bool     front = (flags & TRI_FLIPPED) ? (det < 0.0f) : (det > 0.0f);
        hit &= !cull | ((flags & TRI_NOCULL) != 0) | front;
        hit &= !haveLo | (t > loT) | ((t == loT) & (gid > loId));
        hit &= (t < best.t) | ((t == best.t) & (gid < best.gid));
        if(hit)
        {
          if(collectN >= 0)
          {
            // insertion sort by (t, id); a full list drops its last entry (the predicate above already
            // guarantees the new hit sorts before it)
            int pos = collectN < kCand ? collectN : kCand - 1;
            while(pos > 0 && ((cand[pos - 1].t > t) | ((cand[pos - 1].t == t) & (cand[pos - 1].gid > gid))))
            {
              cand[pos] = cand[pos - 1];
              pos--;
            }
            cand[pos].t = t;
            cand[pos].u = u;
            cand[pos].v = v;
            cand[pos].slot = slot;
            cand[pos].gid = gid;
            if(collectN < kCand)
              collectN++;
            if(collectN == kCand)
            {
              best.t = cand[kCand - 1].t;
              best.gid = cand[kCand - 1].gid;
            }
          }
          else
          {
            best.t = t;
            best.u = u;
            best.v = v;
            best.slot = slot;
            best.gid = gid;
            best.w0 = w0;
            done = anyExit;  // occlusion query satisfied
          }
        }
      }
    }
    return done | ((tri.y == 0) & ((cur.y & 0xff000000u) == 0) & (sp == 0));
  }

  // the hit with (u,v) restored for mirrored instances
  PT_D TraceHit result() const
  {
    TraceHit h = best;
    if(h.slot != 0xFFFFFFFFu && ((h.w0 >> 28) & TRI_FLIPPED))
    {
      h.u = best.v;
      h.v = best.u;
    }
    return h;
  }

  PT_D void flushCounters(unsigned long long* nodeCounter, unsigned long long* triCounter)
  {
#ifdef B200PT_COUNT_TRAVERSAL
    if(nodeCounter)
      atomicAdd(nodeCounter, (unsigned long long)nodeCount);
    if(triCounter)
      atomicAdd(triCounter, (unsigned long long)triCount);
    nodeCount = triCount = 0;
#else
    (void)nodeCounter;
    (void)triCounter;
#endif
  }
};

// one collecting walk to completion: the up-to-kCand nearest hits after the lower bound, sorted, in `cand`
PT_D int collectNext(const BvhView bvh, float3 org, float3 dir, float tmax, bool cull, bool haveLo, float loT, uint32_t loId, Cand* __restrict__ cand)
{
  TravState T;
  uint2     stack[TravState::kStackSize];
  T.init(bvh, org, dir, 0.0f, tmax, cull, false, haveLo, loT, loId, true);
  while(!T.step(stack, 2, cand))
  {
  }
  return T.collectN;
}

// run one traversal to completion (ray-level API kernels and the any-hit restart loops)
template <bool CULL, bool ANY_EXIT>
PT_D TraceHit traverseNext(const BvhView bvh, float3 org, float3 dir, float tmin, float tmax, bool haveLo, float loT, uint32_t loId,
                           unsigned long long* nodeCounter = nullptr, unsigned long long* triCounter = nullptr)
{
  TravState T;
  uint2     stack[TravState::kStackSize];
  T.init(bvh, org, dir, tmin, tmax, CULL, ANY_EXIT, haveLo, loT, loId);
  while(!T.step(stack))
  {
  }
  T.flushCounters(nodeCounter, triCounter);
  return T.result();
}

}  // namespace pt
