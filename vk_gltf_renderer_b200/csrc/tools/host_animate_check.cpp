// host_animate_check.cpp — the device animation source (animate.cuh: morphVertex / skinVertex, the bodies of k_morph / k_skin)
// compiled for the host through host_shim.h and run vertex by vertex on a task dumped by tests/test_animation.py; the outputs are
// written back for a bit-for-bit comparison with oracle/animation.py.
//   host_animate_check task.bin out.bin
// task.bin: u32 kind (0 morph, 1 skin), V, K (targets / joints), hasN, hasT, hasDN, hasDT, then the float arrays in the order of
// MorphTaskDev / SkinTaskDev (joints as int32).  kind 2 = the rigid part (propagateNode level by level, then updateRenderNode):
// V = graph nodes, K = render nodes, hasN = levels, hasT = instance matrices present; arrays: parents i32[V], topo i32[V],
// level offsets u32[levels + 1], mappings 4 x i32 [K], local f32[V x 16], instLocal f32[K x 16]?; out.bin: world f32[V x 16], then K
// b200pt_render_node records.
#include "host_shim.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../animate.cuh"

using namespace pt;

template <typename T>
static std::vector<T> rd(FILE* f, size_t n)
{
  std::vector<T> v(n);
  if(n && std::fread(v.data(), sizeof(T), n, f) != n)
  {
    std::fprintf(stderr, "short read\n");
    std::exit(2);
  }
  return v;
}

int main(int argc, char** argv)
{
  if(argc < 3)
    return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if(!f)
    return 2;
  const std::vector<uint32_t> hd = rd<uint32_t>(f, 7);
  const uint32_t kind = hd[0], V = hd[1], K = hd[2];
  const bool     hasN = hd[3], hasT = hd[4], hasDN = hd[5], hasDT = hd[6];
  if(kind == 2)
  {
    const uint32_t              levels = hd[3];
    const std::vector<int>      parents = rd<int>(f, V), topo = rd<int>(f, V);
    const std::vector<uint32_t> ofs = rd<uint32_t>(f, levels + 1);
    const std::vector<RenderNodeMapping> maps = rd<RenderNodeMapping>(f, K);
    const std::vector<float>    local = rd<float>(f, (size_t)V * 16), inst = rd<float>(f, hd[4] ? (size_t)K * 16 : 0);
    std::fclose(f);
    std::vector<float>              world((size_t)V * 16, 0.f);
    std::vector<b200pt_render_node> nodes(K);
    for(uint32_t l = 0; l < levels; l++)
      for(uint32_t ti = 0; ti < ofs[l + 1] - ofs[l]; ti++)
        propagateNode(local.data(), world.data(), parents.data(), topo.data(), ofs[l], ti);
    for(uint32_t i = 0; i < K; i++)
      updateRenderNode(world.data(), maps.data(), hd[4] ? inst.data() : nullptr, nodes.data(), i);
    FILE* o = std::fopen(argv[2], "wb");
    if(!o)
      return 2;
    std::fwrite(world.data(), 4, world.size(), o);
    std::fwrite(nodes.data(), sizeof(b200pt_render_node), nodes.size(), o);
    std::fclose(o);
    return 0;
  }
  std::vector<float> outP((size_t)V * 3), outN((size_t)V * 3), outT((size_t)V * 4);
  const std::vector<float> bp = rd<float>(f, (size_t)V * 3), bn = rd<float>(f, hasN ? (size_t)V * 3 : 0), bt = rd<float>(f, hasT ? (size_t)V * 4 : 0);
  if(kind == 0)
  {
    const std::vector<float> dp = rd<float>(f, (size_t)K * V * 3), dn = rd<float>(f, hasDN ? (size_t)K * V * 3 : 0), dt = rd<float>(f, hasDT ? (size_t)K * V * 3 : 0);
    const std::vector<float> w = rd<float>(f, K);
    MorphTaskDev T{};
    T.basePos = bp.data(), T.baseNrm = hasN ? bn.data() : nullptr, T.baseTan = hasT ? bt.data() : nullptr;
    T.dPos = dp.data(), T.dNrm = hasDN ? dn.data() : nullptr, T.dTan = hasDT ? dt.data() : nullptr;
    T.weights = w.data();
    T.outPos = outP.data(), T.outNrm = hasN ? outN.data() : nullptr, T.outTan = hasT ? outT.data() : nullptr;
    T.vertexCount = V, T.numTargets = K;
    for(uint32_t v = 0; v < V; v++)
      morphVertex(T, v);
  }
  else
  {
    const std::vector<float> w = rd<float>(f, (size_t)V * 4);
    const std::vector<int>   j = rd<int>(f, (size_t)V * 4);
    const std::vector<float> jm = rd<float>(f, (size_t)K * 16), nm = rd<float>(f, (size_t)K * 9);
    SkinTaskDev T{};
    T.basePos = bp.data(), T.baseNrm = hasN ? bn.data() : nullptr, T.baseTan = hasT ? bt.data() : nullptr;
    T.weights = w.data(), T.joints = j.data(), T.jointMatrices = jm.data(), T.normalMatrices = nm.data();
    T.outPos = outP.data(), T.outNrm = hasN ? outN.data() : nullptr, T.outTan = hasT ? outT.data() : nullptr;
    T.vertexCount = V, T.numJoints = K;
    for(uint32_t v = 0; v < V; v++)
      skinVertex(T, v);
  }
  std::fclose(f);
  FILE* o = std::fopen(argv[2], "wb");
  if(!o)
    return 2;
  std::fwrite(outP.data(), 4, outP.size(), o);
  if(hasN)
    std::fwrite(outN.data(), 4, outN.size(), o);
  if(hasT)
    std::fwrite(outT.data(), 4, outT.size(), o);
  std::fclose(o);
  std::printf("kind %u vertices %u k %u\n", kind, V, K);
  return 0;
}
