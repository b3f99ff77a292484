// simt_sim.cpp — offline SIMT model of the persistent-warp tree walk (k_trace / k_shadow inner loop).
//
// Walks the same compressed 8-wide BVH as the kernel, 32 rays per simulated warp, with the kernel's scheduling rules:
// one node step or one triangle test per lane per iteration, triangle groups postponed when < 1/4 of the lanes have one,
// refill from a shared cursor when fewer than `refill` lanes are busy.  It counts what the GPU pays for: warp iterations
// in which the node phase / the triangle phase is executed at all (every executed phase costs its full instruction
// count regardless of how many lanes take part).  The instruction weights (node 270, triangle 110, loop 25, refill 150)
// come from the SASS of the real kernels.  CAVEAT: it is an issue-slot model only.  It reproduces the measured lane
// statistics roughly (busy 26 vs 21.5 measured incl. launch tails, triangle-phase lanes 10.9 vs 10.6) but it predicts
// variant 2 to be 8 % cheaper while the GPU measured it 20 % slower (two dependent memory round trips and more
// local-memory stack traffic per iteration) -- use it for lane / iteration statistics, not to rank kernels.
//
//   simt_sim dump.bin [refill=22] [postponeShift=2] [variant]
//     variant 0: kernel as shipped            1: + stale-entry culling (min entry distance of the children left at push)
//             2: node AND triangle phase per lane per step ("v2", measured slower on the GPU: the model must agree)
//             3: two triangle tests per triangle phase
//
// Input: the dump written by scripts/dump_bvh_input.py (triangles + mixed primary / diffuse rays).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../bvh.h"
using namespace pt;

static inline uint32_t fu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    uf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct G { uint32_t x, y; float gmin; };

struct Lane
{
  bool     busy = false;
  float    org[3], dir[3], id[3], tmin, best;
  uint32_t octInv;
  G        cur, tri;
  G        stack[48];
  int      sp;
  uint64_t nodes = 0, tris = 0;
};

static const WideBvh* gB;
static int            gVariant = 0;

static void initLane(Lane& L, const float* R)
{
  L.busy = true;
  for(int a = 0; a < 3; a++) { L.org[a] = R[a]; L.dir[a] = R[4 + a]; float d = fabsf(L.dir[a]) > 1e-20f ? L.dir[a] : copysignf(1e-20f, L.dir[a]); L.id[a] = 1.0f / d; }
  L.tmin = R[3]; L.best = R[7];
  L.octInv = (L.dir[0] < 0 ? 0 : 4) | (L.dir[1] < 0 ? 0 : 2) | (L.dir[2] < 0 ? 0 : 1);
  L.cur = G{0, 0x80000000u, -1.f}; L.tri = G{0, 0, 0.f}; L.sp = 0;
}

// node phase of one lane; returns true if a node was opened (i.e. the expensive code ran)
static bool nodePhase(Lane& L)
{
  const WideBvh& B = *gB;
  for(;;)
  {
    if((L.cur.y & 0xff000000u) == 0)
    {
      if(L.sp == 0) return false;
      G e = L.stack[--L.sp];
      if((e.y & 0xff000000u) == 0) { L.tri = e; return false; }  // postponed triangle group
      if(gVariant == 1 && e.gmin > L.best) continue;             // stale: nothing in this group can be nearer any more
      L.cur = e;
    }
    break;
  }
  uint32_t him = L.cur.y; int cb = 31 - __builtin_clz(him); L.cur.y &= ~(1u << cb);
  float tnRemaining[8];
  (void)tnRemaining;
  if(L.cur.y & 0xff000000u) L.stack[L.sp++] = L.cur;  // gmin of this remainder was computed when the node was opened (below)
  uint32_t slot = (uint32_t)(cb - 24) ^ L.octInv;
  uint32_t rel = __builtin_popcount(him & ~(0xffffffffu << slot));
  const float* N = &B.nodes[(size_t)(L.cur.x + rel) * 20];
  L.nodes++;
  uint32_t eim = fu(N[3]);
  float ad[3], ao[3];
  for(int a = 0; a < 3; a++) { ad[a] = uf(((eim >> (8 * a)) & 0xff) << 23) * L.id[a]; ao[a] = (N[a] - L.org[a]) * L.id[a]; }
  uint32_t hm = 0; float tnOf[8]; for(int b = 0; b < 8; b++) tnOf[b] = 1e38f;
  for(int c = 0; c < 8; c++)
  {
    uint32_t meta = (fu(N[6 + c / 4]) >> (8 * (c % 4))) & 0xff;
    bool inner = (meta & (meta << 1)) & 0x10;
    uint32_t bitIndex = (meta ^ (inner ? L.octInv : 0)) & 0x1f, childBits = (meta >> 5) & 7;
    float tn = L.tmin, tf = L.best;
    for(int a = 0; a < 3; a++)
    {
      uint32_t lo = (fu(N[8 + a * 4 + c / 4]) >> (8 * (c % 4))) & 0xff, hi = (fu(N[8 + a * 4 + 2 + c / 4]) >> (8 * (c % 4))) & 0xff;
      float t0 = (L.dir[a] < 0 ? hi : lo) * ad[a] + ao[a], t1 = (L.dir[a] < 0 ? lo : hi) * ad[a] + ao[a];
      tn = fmaxf(tn, t0); tf = fminf(tf, t1);
    }
    if(tn <= tf * 1.000001f) { hm |= childBits << bitIndex; if(inner) tnOf[bitIndex - 24] = tn; }
  }
  L.cur.x = (fu(N[4]) & 0x03ffffffu);
  L.cur.y = (hm & 0xff000000u) | (eim >> 24);
  // gmin for the remainder this group will leave behind once its first child is taken
  {
    uint32_t h = L.cur.y & 0xff000000u;
    float    g = 1e38f;
    if(h) { int first = 31 - __builtin_clz(h); for(int b = 24; b < 32; b++) if((h & (1u << b)) && b != first) g = fminf(g, tnOf[b - 24]); }
    L.cur.gmin = g;
  }
  L.tri = G{fu(N[5]), hm & 0x00ffffffu, 0.f};
  return true;
}

static void triTest(Lane& L)
{
  const WideBvh& B = *gB;
  int tb = 31 - __builtin_clz(L.tri.y); L.tri.y &= ~(1u << tb);
  const float* T = &B.tris[(size_t)(L.tri.x + tb) * 12];
  L.tris++;
  float e1[3] = {T[4], T[5], T[6]}, e2[3] = {T[8], T[9], T[10]};
  float p[3] = {L.dir[1] * e2[2] - L.dir[2] * e2[1], L.dir[2] * e2[0] - L.dir[0] * e2[2], L.dir[0] * e2[1] - L.dir[1] * e2[0]};
  float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
  if(det == 0) return;
  float inv = 1 / det, tv3[3] = {L.org[0] - T[0], L.org[1] - T[1], L.org[2] - T[2]};
  float u = (tv3[0] * p[0] + tv3[1] * p[1] + tv3[2] * p[2]) * inv;
  if(u < 0 || u > 1) return;
  float q[3] = {tv3[1] * e1[2] - tv3[2] * e1[1], tv3[2] * e1[0] - tv3[0] * e1[2], tv3[0] * e1[1] - tv3[1] * e1[0]};
  float v = (L.dir[0] * q[0] + L.dir[1] * q[1] + L.dir[2] * q[2]) * inv;
  if(v < 0 || u + v > 1) return;
  float t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
  if(t > L.tmin && t < L.best) L.best = t;
}

int main(int argc, char** argv)
{
  if(argc < 2) { fprintf(stderr, "usage: simt_sim dump.bin [refill] [postponeShift] [variant]\n"); return 1; }
  const int refill = argc > 2 ? atoi(argv[2]) : 22, pshift = argc > 3 ? atoi(argv[3]) : 2;
  gVariant = argc > 4 ? atoi(argv[4]) : 0;
  FILE* f = fopen(argv[1], "rb");
  uint32_t nT, nR;
  if(!f || fread(&nT, 4, 1, f) != 1 || fread(&nR, 4, 1, f) != 1) return 1;
  std::vector<float> tv((size_t)nT * 9), rv((size_t)nR * 8);
  if(fread(tv.data(), 4, tv.size(), f) != tv.size() || fread(rv.data(), 4, rv.size(), f) != rv.size()) return 1;
  fclose(f);
  std::vector<FlatTri> tris(nT); std::vector<uint32_t> gids(nT);
  for(uint32_t i = 0; i < nT; i++) { memcpy(tris[i].v0, &tv[i * 9], 12); memcpy(tris[i].e1, &tv[i * 9 + 3], 12); memcpy(tris[i].e2, &tv[i * 9 + 6], 12);
    tris[i].rnode = 0; tris[i].prim = i; tris[i].flags = TRI_OPAQUE | TRI_NOCULL; gids[i] = i; }
  WideBvh B; buildWideBvh(tris, gids, 0, B); gB = &B;

  // the GPU runs 148 x 6 x 4 = 3552 resident warps on ~2 M rays (~580 rays per warp); keep that ratio
  const int nWarps = std::max(1, (int)(nR / 580));
  std::vector<std::vector<Lane>> warps(nWarps, std::vector<Lane>(32));
  uint32_t cursor = 0;
  const double cNode = 270, cTri = 110, cLoop = 25, cRefill = 150;
  double cost = 0; uint64_t iters = 0, nodeExec = 0, triExec = 0, busySum = 0, nodeLanes = 0, triLanes = 0, refills = 0;
  std::vector<char> alive(nWarps, 1);
  int nAlive = nWarps;
  while(nAlive)
  {
    for(int w = 0; w < nWarps; w++)
    {
      if(!alive[w]) continue;
      auto& W = warps[w];
      // ---- outer loop: refill ----
      int busy = 0; for(auto& L : W) busy += L.busy;
      if(busy < refill)
      {
        bool any = false;
        for(auto& L : W) if(!L.busy && cursor < nR) { initLane(L, &rv[(size_t)cursor++ * 8]); any = true; }
        if(any) { cost += cRefill; refills++; }
        busy = 0; for(auto& L : W) busy += L.busy;
        if(busy == 0) { alive[w] = 0; nAlive--; continue; }
      }
      // ---- one inner iteration ----
      iters++; busySum += busy; cost += cLoop;
      bool nodeRan = false; int nl = 0;
      for(auto& L : W)
        if(L.busy && (L.tri.y == 0 || gVariant == 2))
        {
          if(gVariant == 2 && L.tri.y != 0)
          {
            // v2: lanes with pending triangles still open a node when they have one; a new leaf group is postponed
            G keep = L.tri;
            if((L.cur.y & 0xff000000u) || (L.sp > 0 && (L.stack[L.sp - 1].y & 0xff000000u)))
            {
              bool ran = nodePhase(L);
              if(ran) { nodeRan = true; nl++; if(L.tri.y) L.stack[L.sp++] = L.tri; }
              L.tri = keep;
            }
            continue;
          }
          bool ran = nodePhase(L);
          if(ran) { nodeRan = true; nl++; }
          if(!ran && L.tri.y == 0 && (L.cur.y & 0xff000000u) == 0 && L.sp == 0) L.busy = false;
        }
      if(nodeRan) { cost += cNode; nodeExec++; nodeLanes += nl; }
      int conv = 0, have = 0;
      for(auto& L : W) if(L.busy) { conv++; if(L.tri.y) have++; }
      bool triRan = false; int tl = 0;
      for(auto& L : W)
        if(L.busy && L.tri.y)
        {
          if(gVariant != 2 && ((have << pshift) < conv) && (L.cur.y & 0xff000000u) != 0 && L.sp < 40) { L.stack[L.sp++] = L.tri; L.tri.y = 0; }
          else { triTest(L); if(gVariant == 3 && L.tri.y) triTest(L); triRan = true; tl++; }
        }
      if(triRan) { cost += (gVariant == 3 ? 2 * cTri * 0.9 : cTri); triExec++; triLanes += tl; }
      for(auto& L : W) if(L.busy && L.tri.y == 0 && (L.cur.y & 0xff000000u) == 0 && L.sp == 0) L.busy = false;
    }
  }
  uint64_t tn = 0, tt = 0; for(auto& W : warps) for(auto& L : W) { tn += L.nodes; tt += L.tris; }
  printf("variant %d refill %d postpone %d: %.2f nodes/ray %.2f tris/ray | %.1f warp-instr/ray | iterations %.2f/ray, busy %.1f/32, node phase in %.0f%% of iterations (%.1f lanes), triangle phase in %.0f%% (%.1f lanes), refills %.3f/ray\n",
         gVariant, refill, pshift, (double)tn / nR, (double)tt / nR, cost / nR, (double)iters / nR, (double)busySum / iters, 100.0 * nodeExec / iters,
         (double)nodeLanes / std::max<uint64_t>(nodeExec, 1), 100.0 * triExec / iters, (double)triLanes / std::max<uint64_t>(triExec, 1), (double)refills / nR);
  return 0;
}
