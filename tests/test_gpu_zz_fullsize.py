"""-m gpu, BASELINE.json's full size (config 3's shape: 1920x1080, depth 12, the bench workload): the oracle needs minutes per
frame there, so the checks are size-independent properties of the path -- determinism per (pixel, frame) seed, well-formed
output, ray budget, the running-mean accumulation identity, tile reassembly (multi-GPU partition) and frames in flight."""
import argparse
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, DEPTH = 1920, 1080, 12


@pytest.fixture(scope="module")
def workload():
    sys.path.insert(0, ROOT)
    import bench
    return bench.build_workload(argparse.Namespace(tex=2048, detail=1.0))


def _tracer(scn, env, tile=None, lanes=None):
    from vk_gltf_renderer_b200.renderer import PathTracer, Resources
    res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(W, H), tile=tile)
    pt = PathTracer(0)
    pt.ptMaxDepth = DEPTH
    pt.onAttach(res)
    if lanes:
        pt.set_frames_in_flight(lanes)
    return pt, res


def _frames(pt, res, n):
    res.frameCount = -1
    for _ in range(n):
        res.frameCount += 1
        pt.onRender(None, res)
    return pt.read_accum()


def test_fullsize_deterministic_wellformed_and_ray_budget(workload):
    scn, env = workload
    pt, res = _tracer(scn, env)
    pt.reset_stats()
    a = _frames(pt, res, 2)
    st = pt.stats()
    assert a.shape == (H, W, 4) and np.isfinite(a).all() and (a[..., :3] >= 0).all()
    assert (a[..., 3] >= 0).all() and (a[..., 3] <= 1).all()  # .w: running mean of the primary hit's solid flag
    # two frames of a 0/1 flag average to 0, 0.5 or 1 -- except where a frame's sample was firefly-clamped: the clamp scales all
    # four channels (gltf_pathtrace.slang:533-538), so such a pixel carries a fractional .w
    assert np.isin(a[..., 3], (0.0, 0.5, 1.0)).mean() > 0.9
    assert 0.02 < float(a[..., :3].mean()) < 50.0
    paths = 2 * W * H
    assert st["pathsStarted"] == paths
    assert paths <= st["closestRays"] <= paths * DEPTH      # at least the camera ray, at most one per bounce
    assert st["shadowRays"] <= st["shadedHits"] <= st["closestRays"]
    pt.onDetach()
    pt2, res2 = _tracer(scn, env)
    b = _frames(pt2, res2, 2)
    assert np.array_equal(a, b)                             # same seeds, same bits: nothing depends on scheduling


def test_fullsize_running_mean_identity(workload):
    """processPixel's accumulation (old * total + new * n) / (total + n): two frames accumulate to the mean of the two
    frames rendered on their own (frame 1 alone = frameCount 1 with the first-frame flag forcing an overwrite)."""
    from vk_gltf_renderer_b200 import abi, camera as cm
    scn, env = workload
    pt, res = _tracer(scn, env)
    both = _frames(pt, res, 2)
    f0 = _frames(pt, res, 1)
    fi = cm.make_frame_info(scn.camera, W, H)
    pc = cm.make_push_constant(scn.camera, H, frame_count=1, total_samples=0, max_depth=DEPTH)
    pc.flags = abi.PT_FIRST_FRAME
    pt.render_frame_raw(fi, pc)
    f1 = pt.read_accum()
    assert not np.array_equal(f0, f1)
    expect = (f0 * np.float32(1.0) + f1 * np.float32(1.0)) / np.float32(2.0)
    assert np.allclose(both, expect, rtol=2e-6, atol=1e-7)


def test_fullsize_tiles_reassemble_the_frame(workload):
    """The 8-rank partition of bench.py (interleaved bands) rendered rank by rank on one GPU reassembles to the full frame
    bit for bit, and so does a contiguous strip."""
    from vk_gltf_renderer_b200 import tiling
    scn, env = workload
    pt, res = _tracer(scn, env)
    full = _frames(pt, res, 2)
    world = 8
    band = tiling.interleave_band(H, world)
    parts = []
    for r in range(world):
        res.tile = ("interleave", band, world, r)
        pt.onResize(None, (W, H), res)
        parts.append(_frames(pt, res, 2))
        assert parts[-1].shape == (H // world, W, 4)
    assert np.array_equal(tiling.deinterleave(np.concatenate(parts), H, world, band), full)
    res.tile = (400, 135)
    pt.onResize(None, (W, H), res)
    assert np.array_equal(_frames(pt, res, 2), full[400:535])


def test_fullsize_frames_in_flight_do_not_change_the_image(workload):
    scn, env = workload
    imgs = []
    for lanes in (1, 4):
        pt, res = _tracer(scn, env, lanes=lanes)
        imgs.append(_frames(pt, res, 5))
        pt.onDetach()
    assert np.array_equal(imgs[0], imgs[1])
