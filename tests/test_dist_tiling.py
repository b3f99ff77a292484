"""CPU, world_size 2 over gloo: the N>1 path of bench.py — row-strip partition, per-rank rendering with GLOBAL
pixel coordinates in the seed, one all-gather per frame — reproduces the single-rank image bit for bit.
(The per-rank renderer here is the CPU oracle: no GPU in this container; on the B200 box the same partition and
gather code runs over NCCL with the CUDA renderer, see bench.py.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker_interleaved(rank, world, port, out_path):
    """bench.py's N>1 partition: interleaved bands, all-gather in rank-major order, de-interleave."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from vk_gltf_renderer_b200 import camera as cm, hdr, scene, tiling
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scn = scene.load_gltf(os.path.join(ROOT, "tests", "assets", "Box.glb"))
    env = hdr.load_hdr(os.path.join(ROOT, "tests", "assets", "std_env.hdr"))
    o = O.Oracle()
    o.set_scene(scn)
    o.set_environment(env)
    W, H, frames = 48, 36, 2
    band = tiling.interleave_band(H, world, max_band=3)
    assert band == 3
    rows = tiling.interleaved_rows(H, world, rank, band)
    tile = np.zeros((H // world, W, 4), np.float32)
    full = torch.empty((H, W, 4), dtype=torch.float32)
    fi = cm.make_frame_info(scn.camera, W, H)
    for f in range(frames):
        pc = cm.make_push_constant(scn.camera, H, frame_count=f, total_samples=f, max_depth=4)
        for b in range(0, len(rows), band):  # the oracle renders contiguous strips: one call per owned band
            o.render_frame(fi, pc, tile[b:b + band], y0=rows[b], rows=band, threads=1)
        dist.all_gather_into_tensor(full, torch.from_numpy(tile))
    if rank == 0:
        np.save(out_path, tiling.deinterleave(full, H, world, band).numpy())
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from vk_gltf_renderer_b200 import camera as cm, hdr, scene, tiling
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scn = scene.load_gltf(os.path.join(ROOT, "tests", "assets", "Box.glb"))
    env = hdr.load_hdr(os.path.join(ROOT, "tests", "assets", "std_env.hdr"))
    o = O.Oracle()
    o.set_scene(scn)
    o.set_environment(env)
    W, H, frames = 48, 37, 3  # odd height: the last strip is shorter than the padded strip
    per = tiling.strip_rows(H, world)
    y0, rows = tiling.partition_rows(H, world, rank)
    tile = np.zeros((per, W, 4), np.float32)
    full = torch.empty((world * per, W, 4), dtype=torch.float32)
    fi = cm.make_frame_info(scn.camera, W, H)
    total = 0
    for f in range(frames):
        pc = cm.make_push_constant(scn.camera, H, frame_count=f, total_samples=total, max_depth=4)
        o.render_frame(fi, pc, tile, y0=y0, rows=rows, threads=1)
        total += 1
        dist.all_gather_into_tensor(full, torch.from_numpy(tile))
    if rank == 0:
        np.save(out_path, tiling.assemble(full.numpy(), H))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tiles_equal_single_rank(tmp_path, oracle_mod, box_scene, std_env):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    o.set_environment(std_env)
    ref = oracle_mod.render(o, box_scene.camera, 48, 37, 3, max_depth=4)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_two_rank_interleaved_bands_equal_single_rank(tmp_path, oracle_mod, box_scene, std_env):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered_il.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_interleaved, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    o.set_environment(std_env)
    ref = oracle_mod.render(o, box_scene.camera, 48, 36, 2, max_depth=4)
    assert got.shape == ref.shape and np.array_equal(got, ref)
