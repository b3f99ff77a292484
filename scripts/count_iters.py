"""SIMT bookkeeping of the traversal loops on the bench workload (counter build; needs a GPU).
usage: B200PT_LIB=<path to a counter build> python scripts/count_iters.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200PT_DUMP_ITERS"] = "1"
import bench
from vk_gltf_renderer_b200.renderer import PathTracer, Resources
sys.argv = sys.argv[:1]
args = bench.parse()
scn, env = bench.build_workload(args)
res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(args.width, args.height))
pt = PathTracer(0)
pt.ptMaxDepth = args.depth
pt.onAttach(res)
res.frameCount = 0
pt.onRender(None, res)
st = pt.stats()
print({k: st[k] for k in ("closestRays", "shadowRays", "nodesVisited", "trisTested")})
