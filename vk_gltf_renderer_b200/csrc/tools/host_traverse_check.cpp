// host_traverse_check.cpp — the device traversal source (traverse.cuh: node decode with the PRMT plane conversion, octant
// order, compressed stack, postponed groups, nearest-hit and collecting modes, lower bounds) compiled for the host through
// host_shim.h and checked against brute force over all triangles, on the tree the product's builder (bvh.cpp) makes.
//
//   host_traverse_check dump.bin        (dump format of scripts/dump_bvh_input.py: triangles + rays)
//
// Checks, per ray: (1) nearest hit: same (t, global id) as the brute-force minimum in (t, id) order, bit for bit (the triangle
// test is the same fma chain); (2) any-exit query: hits iff any triangle hits; (3) collecting mode: the kCand nearest hits in
// (t, id) order equal the head of the sorted brute-force list, and walking on with the last one as lower bound enumerates the
// whole list in order -- the sequence the any-hit kernels consume.  Exit code 0 iff everything matches.
#include "host_shim.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../traverse.cuh"

using namespace pt;

struct BF
{
  float    t;
  uint32_t gid;
};

int main(int argc, char** argv)
{
  if(argc < 2) { std::fprintf(stderr, "usage: host_traverse_check dump.bin [maxRays]\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  uint32_t nT, nR;
  if(!f || std::fread(&nT, 4, 1, f) != 1 || std::fread(&nR, 4, 1, f) != 1) return 2;
  std::vector<float> tv((size_t)nT * 9), rv((size_t)nR * 8);
  if(std::fread(tv.data(), 4, tv.size(), f) != tv.size() || std::fread(rv.data(), 4, rv.size(), f) != rv.size()) return 2;
  std::fclose(f);
  if(argc > 2) nR = std::min<uint32_t>(nR, (uint32_t)std::atoi(argv[2]));
  std::vector<FlatTri>  tris(nT);
  std::vector<uint32_t> gids(nT);
  for(uint32_t i = 0; i < nT; i++)
  {
    std::memcpy(tris[i].v0, &tv[i * 9], 12); std::memcpy(tris[i].e1, &tv[i * 9 + 3], 12); std::memcpy(tris[i].e2, &tv[i * 9 + 6], 12);
    tris[i].rnode = 0; tris[i].prim = i; tris[i].flags = TRI_NOCULL; gids[i] = i;
  }
  WideBvh B;
  buildWideBvh(tris, gids, 0, B);
  BvhView view{reinterpret_cast<const float4*>(B.nodes.data()), reinterpret_cast<const float4*>(B.tris.data())};
  std::printf("tris %u nodes %u rays %u kCand %d\n", B.numTris, B.numNodes, nR, kCand);

  uint64_t bad1 = 0, bad2 = 0, bad3 = 0, hits = 0, listed = 0;
  int      maxSp = 0;
  std::vector<BF> bf;
  for(uint32_t r = 0; r < nR; r++)
  {
    const float* R = &rv[(size_t)r * 8];
    const float3 org = f3(R[0], R[1], R[2]), dir = f3(R[4], R[5], R[6]);
    const float  tmin = 0.0f, tmax = (r & 1) ? R[7] : 3.0f;  // half the rays as bounded segments
    // brute force with the kernel's own triangle arithmetic
    bf.clear();
    for(uint32_t i = 0; i < nT; i++)
    {
      const float3 v0 = f3(tris[i].v0[0], tris[i].v0[1], tris[i].v0[2]), e1 = f3(tris[i].e1[0], tris[i].e1[1], tris[i].e1[2]),
                   e2 = f3(tris[i].e2[0], tris[i].e2[1], tris[i].e2[2]);
      const float3 pvec = crossFma(dir, e2);
      const float  det = dotFma(e1, pvec);
      const float  inv = 1.0f / det;
      const float3 tvec = org - v0;
      const float  u = dotFma(tvec, pvec) * inv;
      const float3 qvec = crossFma(tvec, e1);
      const float  v = dotFma(dir, qvec) * inv;
      const float  t = dotFma(e2, qvec) * inv;
      if((det != 0.0f) & (u >= 0.0f) & (u <= 1.0f) & (v >= 0.0f) & (u + v <= 1.0f) & (t > tmin) & (t < tmax))
        bf.push_back(BF{t, i});
    }
    std::sort(bf.begin(), bf.end(), [](const BF& a, const BF& b) { return a.t < b.t || (a.t == b.t && a.gid < b.gid); });
    // (1) nearest hit (stepped by hand to record the deepest stack the walk needed; kStackSize entries exist)
    TraceHit h;
    {
      TravState T;
      uint2     stack[TravState::kStackSize];
      T.init(view, org, dir, tmin, tmax, false, false, false, 0.f, 0u);
      while(!T.step(stack))
        maxSp = std::max(maxSp, T.sp);
      h = T.result();
    }
    if(bf.empty() ? (h.slot != 0xFFFFFFFFu) : (h.slot == 0xFFFFFFFFu || __float_as_uint(h.t) != __float_as_uint(bf[0].t) || h.gid != bf[0].gid))
      bad1++;
    hits += !bf.empty();
    // (2) any-exit occlusion query
    const TraceHit a = traverseNext<false, true>(view, org, dir, tmin, tmax, false, 0.f, 0u);
    if((a.slot != 0xFFFFFFFFu) != !bf.empty())
      bad2++;
    // (3) collecting walks, kCand at a time, resumed behind the last candidate
    {
      Cand     cand[kCand];
      size_t   k = 0;
      bool     haveLo = false, ok = true;
      float    loT = 0.f;
      uint32_t loId = 0;
      for(;;)
      {
        const int m = collectNext(view, org, dir, tmax, false, haveLo, loT, loId, cand);
        for(int i = 0; i < m && ok; i++, k++)
          ok = k < bf.size() && __float_as_uint(cand[i].t) == __float_as_uint(bf[k].t) && cand[i].gid == bf[k].gid;
        if(!ok || m < kCand)
          break;
        haveLo = true;
        loT = cand[kCand - 1].t;
        loId = cand[kCand - 1].gid;
      }
      if(!ok || k != bf.size())
        bad3++;
      listed += k;
    }
  }
  std::printf("hit rate %.3f, %.2f candidates per ray, deepest stack %d of %d | mismatches: nearest %llu, any-exit %llu, collecting %llu\n", (double)hits / nR,
              (double)listed / nR, maxSp, TravState::kStackSize, (unsigned long long)bad1, (unsigned long long)bad2, (unsigned long long)bad3);
  if(maxSp >= TravState::kStackSize)
    return 1;
  return (bad1 | bad2 | bad3) ? 1 : 0;
}
