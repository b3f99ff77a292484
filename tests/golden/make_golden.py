"""Regenerates the golden fixtures in this directory from the CPU oracle (oracle/, the restatement of the reference
algorithm -- the reference itself cannot be built or run offline, DESIGN.md section 4, so these vectors pin the ORACLE
and through it the CUDA path; they are not outputs of the Vulkan reference).

    python tests/golden/make_golden.py          # rewrites the .npz files next to this script

Fixtures (all small, seeds fixed):
  box_64x64_f2_d4.npz      Box.glb + std_env.hdr, 64x64, 2 frames x 1 spp, depth 4: RGBA32F accumulation image
  layers_mask_48x36.npz    14 alpha-MASK layers (any-hit continuation paths), 48x36, 2 frames, depth 4
  soup_rays_2k.npz         2048 rays vs a 3000-triangle soup with a MASK material: closest hits (t,u,v,ids) + seeds after,
                           shadow transmissions + seeds after
  bsdf_256.npz             256 random material / direction records: bsdfEvaluate and bsdfSample outputs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def soup_scene():
    from vk_gltf_renderer_b200 import synth
    scn = synth.triangle_soup(3000, seed=77)
    rng = np.random.default_rng(5)
    a = (rng.random((32, 32)) > 0.45).astype(np.uint8) * 255
    rgba = np.stack([np.full_like(a, 200), np.full_like(a, 180), np.full_like(a, 160), a], -1)
    tex = scn.add_texture(rgba, srgb=True)
    m = scn.materials[0]
    m.alphaMode, m.alphaCutoff, m.doubleSided = 1, 0.5, 1
    m.pbrBaseColorTexture = scn.add_texture_info(tex, 0)
    p = scn.render_prims[0]
    p["uv0"] = np.random.default_rng(9).random((len(p["positions"]), 2)).astype(np.float32)
    return scn


def soup_rays():
    from gpu_util import random_rays
    rays = random_rays(2048, [-1, -1, -1], [1, 1, 1], seed=31)
    seeds = ((np.arange(len(rays), dtype=np.uint64) * 2654435761 + 12345) % (2 ** 32)).astype(np.uint32)
    shadow = rays.copy()
    shadow[:, 7] = 2.5
    return rays, shadow, seeds


def main():
    from oracle import oracle as O
    from vk_gltf_renderer_b200 import bsdf_io, hdr, scene, synth
    env = hdr.load_hdr(os.path.join(ROOT, "tests", "assets", "std_env.hdr"))

    box = scene.load_gltf(os.path.join(ROOT, "tests", "assets", "Box.glb"))
    o = O.Oracle(); o.set_scene(box); o.set_environment(env)
    np.savez_compressed(os.path.join(HERE, "box_64x64_f2_d4.npz"), image=O.render(o, box.camera, 64, 64, 2, max_depth=4))

    lay = synth.synth_layers()
    o = O.Oracle(); o.set_scene(lay); o.set_environment(env)
    np.savez_compressed(os.path.join(HERE, "layers_mask_48x36.npz"), image=O.render(o, lay.camera, 48, 36, 2, max_depth=4))

    scn = soup_scene()
    o = O.Oracle(); o.set_scene(scn)
    rays, shadow, seeds = soup_rays()
    s1 = seeds.copy()
    hits = o.trace_closest(rays, s1)
    s2 = seeds.copy()
    trans = o.trace_shadow(shadow, s2)
    np.savez_compressed(os.path.join(HERE, "soup_rays_2k.npz"), hits=hits, seeds_after_closest=s1, transmission=trans, seeds_after_shadow=s2)

    rec = bsdf_io.random_records(256, seed=4242)
    np.savez_compressed(os.path.join(HERE, "bsdf_256.npz"), records=rec, eval=o.bsdf_eval(rec), sample=o.bsdf_sample(rec))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
