"""glTF / GLB (+ optional Radiance .hdr) -> B2SC scene blob for the C++ host (host/b200pt_host.cpp SceneData::load).

    python -m vk_gltf_renderer_b200.convert scene.glb out.b2sc [--hdr env.hdr]

b200pt_headless calls this on the fly when `--scenefile` names a .gltf / .glb, so that the reference's benchmark harness
(utils/benchmark/benchmark_runner.py:164-198) can hand it the same asset paths it hands vk_gltf_renderer.  The blob holds
exactly what the reference's own loader produces for the path tracer (SceneVk / MaterialCache arrays + decoded textures).
"""
import argparse
import sys

from . import hdr, scene


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("gltf")
    ap.add_argument("out")
    ap.add_argument("--hdr", default=None)
    a = ap.parse_args(argv)
    scn = scene.load_gltf(a.gltf)
    env = hdr.load_hdr(a.hdr) if a.hdr else None
    scn.save_blob(a.out, env)
    print("B2SC %s: %d render nodes, %d triangles, %d materials, %d textures%s" % (a.out, len(scn.render_nodes), scn.num_triangles(), len(scn.materials),
                                                                                  len(scn.textures), ", env %dx%d" % (env.shape[1], env.shape[0]) if env is not None else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
