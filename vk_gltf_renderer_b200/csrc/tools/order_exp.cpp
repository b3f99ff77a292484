// order_exp.cpp — offline experiment: per-node choice of WHICH AXIS each of the 3 slot bits follows.
// Reads the same dump as bvh_stats, builds the wide BVH, then walks it with (a) octant order (b) per-node axis-mapped order
// (c) true front-to-back order, counting node visits.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <array>
#include "../bvh.h"
using namespace pt;
static inline uint32_t fu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    uf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct Child { bool inner; uint32_t node, triBase, triBits; float lo[3], hi[3]; int code; };
struct NodeInfo { std::vector<Child> ch; int axisOfBit[3]; };

int main(int argc, char** argv)
{
  FILE* f = fopen(argv[1], "rb");
  uint32_t nT, nR;
  if(fread(&nT, 4, 1, f) != 1 || fread(&nR, 4, 1, f) != 1) return 1;
  std::vector<float> tv((size_t)nT * 9), rv((size_t)nR * 8);
  if(fread(tv.data(), 4, tv.size(), f) != tv.size() || fread(rv.data(), 4, rv.size(), f) != rv.size()) return 1;
  fclose(f);
  std::vector<FlatTri> tris(nT); std::vector<uint32_t> gids(nT);
  for(uint32_t i = 0; i < nT; i++) { memcpy(tris[i].v0, &tv[i * 9], 12); memcpy(tris[i].e1, &tv[i * 9 + 3], 12); memcpy(tris[i].e2, &tv[i * 9 + 6], 12);
    tris[i].rnode = 0; tris[i].prim = i; tris[i].flags = TRI_OPAQUE | TRI_NOCULL; gids[i] = i; }
  WideBvh B; buildWideBvh(tris, gids, 0, B);
  const uint32_t NN = B.numNodes;
  std::vector<NodeInfo> info(NN);
  // sample directions for the ordering score
  std::vector<std::array<float,3>> dirs;
  for(int x = -1; x <= 1; x++) for(int y = -1; y <= 1; y++) for(int z = -1; z <= 1; z++) if(x || y || z) { float l = std::sqrt((float)(x*x+y*y+z*z)); dirs.push_back({x/l, y/l, z/l}); }
  long changed = 0;
  for(uint32_t n = 0; n < NN; n++)
  {
    const float* N = &B.nodes[(size_t)n * 20];
    uint32_t eim = fu(N[3]), imask = eim >> 24;
    float sc[3]; for(int a = 0; a < 3; a++) sc[a] = uf(((eim >> (8 * a)) & 0xff) << 23);
    NodeInfo& I = info[n];
    for(int c = 0; c < 8; c++)
    {
      uint32_t meta = (fu(N[6 + c / 4]) >> (8 * (c % 4))) & 0xff;
      if(meta == 0) continue;
      Child k{}; k.inner = (meta & (meta << 1)) & 0x10;
      for(int a = 0; a < 3; a++) { uint32_t lo = (fu(N[8 + a * 4 + c / 4]) >> (8 * (c % 4))) & 0xff, hi = (fu(N[8 + a * 4 + 2 + c / 4]) >> (8 * (c % 4))) & 0xff;
        k.lo[a] = N[a] + lo * sc[a]; k.hi[a] = N[a] + hi * sc[a]; }
      if(k.inner) k.node = (fu(N[4]) & 0x03ffffffu) + (uint32_t)__builtin_popcount(imask & ((1u << c) - 1u)); else { k.triBase = fu(N[5]) + (meta & 0x1f); k.triBits = (meta >> 5) & 7; }
      k.code = c;  // octant scheme: slot index
      I.ch.push_back(k);
    }
    // choose mapping: try all 27 axis-of-bit maps, assignment by DP on weighted centroid cost, score by pair-order agreement
    const int nch = (int)I.ch.size();
    float cen[8][3]; float mid[3] = {0,0,0};
    for(int i = 0; i < nch; i++) for(int a = 0; a < 3; a++) { cen[i][a] = 0.5f * (I.ch[i].lo[a] + I.ch[i].hi[a]); }
    for(int a = 0; a < 3; a++) { float lo = 1e30f, hi = -1e30f; for(int i = 0; i < nch; i++) { lo = std::min(lo, cen[i][a]); hi = std::max(hi, cen[i][a]); } mid[a] = 0.5f * (lo + hi); }
    double bestScore = -1; int bestMap = 0; int bestCode[8];
    for(int map = 0; map < 27; map++)
    {
      int ax[3] = {map % 3, (map / 3) % 3, map / 9};  // axis of bit 0,1,2
      // weight of each bit within its axis group: bits on the same axis form a binary rank, higher bit index = more significant
      float w[3]; for(int b = 0; b < 3; b++) { int lower = 0; for(int b2 = 0; b2 < b; b2++) if(ax[b2] == ax[b]) lower++; w[b] = (float)(1 << lower); }
      float cost[8][8];
      for(int i = 0; i < nch; i++) for(int s = 0; s < 8; s++) { float v = 0; for(int b = 0; b < 3; b++) v += ((s >> b) & 1 ? 1.f : -1.f) * w[b] * (cen[i][ax[b]] - mid[ax[b]]); cost[i][s] = v; }
      float best[256]; uint8_t from[8][256]; for(int m = 0; m < 256; m++) best[m] = -1e30f; best[0] = 0;
      for(int m = 0; m < 256; m++) { int i = __builtin_popcount(m); if(i >= nch || best[m] < -1e29f) continue;
        for(int sl = 0; sl < 8; sl++) { if(m & (1 << sl)) continue; int nm = m | (1 << sl); float v = best[m] + cost[i][sl]; if(v > best[nm]) { best[nm] = v; from[i][nm] = sl; } } }
      int bm = -1; float bv = -1e30f; for(int m = 0; m < 256; m++) if(__builtin_popcount(m) == nch && best[m] > bv) { bv = best[m]; bm = m; }
      int code[8]; { int m = bm; for(int i = nch - 1; i >= 0; i--) { int sl = from[i][m]; code[i] = sl; m &= ~(1 << sl); } }
      // score: pair-order agreement over sample directions (entry order ~ centroid . d)
      double score = 0;
      for(auto& d : dirs)
      {
        int eff = 0; for(int b = 0; b < 3; b++) if(d[ax[b]] >= 0) eff |= 1 << b;
        for(int i = 0; i < nch; i++) for(int j = i + 1; j < nch; j++)
        {
          float pi = cen[i][0]*d[0] + cen[i][1]*d[1] + cen[i][2]*d[2], pj = cen[j][0]*d[0] + cen[j][1]*d[1] + cen[j][2]*d[2];
          if(pi == pj) { score += 0.5; continue; }
          // visited first = larger (code ^ eff); nearer child has smaller projection
          bool iFirst = (code[i] ^ eff) > (code[j] ^ eff);
          score += ((pi < pj) == iFirst) ? 1.0 : 0.0;
        }
      }
      if(map == 21) score *= 1.0000001;  // prefer the standard map (bit0=z? see below) on ties
      if(score > bestScore) { bestScore = score; bestMap = map; memcpy(bestCode, code, sizeof(code)); }
    }
    I.axisOfBit[0] = bestMap % 3; I.axisOfBit[1] = (bestMap / 3) % 3; I.axisOfBit[2] = bestMap / 9;
    for(int i = 0; i < nch; i++) I.ch[i].code = bestCode[i];
    if(!(I.axisOfBit[0] == 2 && I.axisOfBit[1] == 1 && I.axisOfBit[2] == 0)) changed++;
  }
  printf("nodes %u, %ld use a non-standard bit->axis map\n", NN, changed);

  auto walk = [&](int mode) {  // 0: per-node mapped code order, 1: distance order
    double nodes = 0, tt = 0;
    for(uint32_t r = 0; r < nR; r++)
    {
      const float* R = &rv[r * 8];
      float org[3] = {R[0], R[1], R[2]}, dir[3] = {R[4], R[5], R[6]}, tmin = R[3], best = R[7], id[3];
      for(int a = 0; a < 3; a++) { float d = fabsf(dir[a]) > 1e-20f ? dir[a] : copysignf(1e-20f, dir[a]); id[a] = 1.0f / d; }
      struct E { uint32_t node; float tn; };
      E stack[256]; int sp = 0; stack[sp++] = E{0, 0.f};
      while(sp)
      {
        E e = stack[--sp]; if(e.tn > best) continue;
        NodeInfo& I = info[e.node]; nodes++;
        int eff = 0; for(int b = 0; b < 3; b++) if(dir[I.axisOfBit[b]] >= 0) eff |= 1 << b;
        struct K { int key; float tn; uint32_t node; }; K kids[8]; int nk = 0;
        for(auto& c : I.ch)
        {
          float tn = tmin, tf = best;
          for(int a = 0; a < 3; a++) { float t0 = ((dir[a] < 0 ? c.hi[a] : c.lo[a]) - org[a]) * id[a], t1 = ((dir[a] < 0 ? c.lo[a] : c.hi[a]) - org[a]) * id[a]; tn = fmaxf(tn, t0); tf = fminf(tf, t1); }
          if(!(tn <= tf * 1.000001f)) continue;
          if(!c.inner)
          {  // leaves of a node are tested right away (as the kernel does)
            uint32_t bits = c.triBits;
            while(bits) { int tb = 31 - __builtin_clz(bits); bits &= ~(1u << tb); const float* T = &B.tris[(size_t)(c.triBase + tb) * 12]; tt++;
              float e1[3] = {T[4], T[5], T[6]}, e2[3] = {T[8], T[9], T[10]};
              float p[3] = {dir[1] * e2[2] - dir[2] * e2[1], dir[2] * e2[0] - dir[0] * e2[2], dir[0] * e2[1] - dir[1] * e2[0]};
              float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2]; if(det == 0) continue;
              float inv = 1 / det, tv3[3] = {org[0] - T[0], org[1] - T[1], org[2] - T[2]};
              float u = (tv3[0] * p[0] + tv3[1] * p[1] + tv3[2] * p[2]) * inv; if(u < 0 || u > 1) continue;
              float q[3] = {tv3[1] * e1[2] - tv3[2] * e1[1], tv3[2] * e1[0] - tv3[0] * e1[2], tv3[0] * e1[1] - tv3[1] * e1[0]};
              float v = (dir[0] * q[0] + dir[1] * q[1] + dir[2] * q[2]) * inv; if(v < 0 || u + v > 1) continue;
              float t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv; if(t > tmin && t < best) best = t; }
            continue;
          }
          kids[nk++] = K{c.code ^ eff, tn, c.node};
        }
        if(mode == 0) std::sort(kids, kids + nk, [](const K& a, const K& b) { return a.key < b.key; });   // push low keys first -> highest key popped first
        else std::sort(kids, kids + nk, [](const K& a, const K& b) { return a.tn > b.tn; });
        for(int i = 0; i < nk; i++) stack[sp++] = E{kids[i].node, kids[i].tn};
      }
    }
    printf("mode %d: %.2f nodes/ray, %.2f tris/ray\n", mode, nodes / nR, tt / nR);
  };
  walk(0); walk(1);
  // and the plain octant scheme through the same walker: reset codes/maps
  for(uint32_t n = 0; n < NN; n++) { info[n].axisOfBit[0] = 2; info[n].axisOfBit[1] = 1; info[n].axisOfBit[2] = 0; }
  // restore slot codes
  for(uint32_t n = 0; n < NN; n++) { const float* N = &B.nodes[(size_t)n * 20]; int k = 0; for(int c = 0; c < 8; c++) { uint32_t meta = (fu(N[6 + c / 4]) >> (8 * (c % 4))) & 0xff; if(meta == 0) continue; info[n].ch[k++].code = c; } }
  printf("octant scheme: "); walk(0);
  return 0;
}
