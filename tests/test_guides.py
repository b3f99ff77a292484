"""Denoiser guide image (OutputImage::eOptixAlbedoNormal, shaders/gltf_pathtrace.slang:240-263, 653-670) in the oracle, without a
GPU: known answers on Box.glb -- primary misses carry albedo 0 and the forward normal (0, 0, 1), hits carry the material's base
colour as a binary16 value and a unit camera-space normal facing the camera, the guide follows the LAST sample of the frame."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def decode_unit_vec(p):
    """inverse of the octahedral 2 x 16-bit encoding (compressUnitVec)"""
    p = np.asarray(p, np.uint32)
    x = (p & 0xFFFF).astype(np.int64) - 32767
    y = (p >> 16).astype(np.int64) - 32767
    ax, ay = np.abs(x), np.abs(y)
    lower = ax + ay > 32767
    fx = np.where(lower, np.sign(x) * (32767 - ay), x).astype(np.float64)
    fy = np.where(lower, np.sign(y) * (32767 - ax), y).astype(np.float64)
    z = (32767 - np.abs(fx) - np.abs(fy)) * np.where(lower, -1.0, 1.0)
    v = np.stack([fx, fy, z], -1)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def render_with_guide(oracle_mod, o, cam, w, h, frames, num_samples=1, max_depth=4):
    from vk_gltf_renderer_b200 import abi, camera as camm
    accum = np.zeros((h, w, 4), np.float32)
    guide = np.zeros((h, w, 4), np.float32)
    fi = camm.make_frame_info(cam, w, h)
    total = 0
    for f in range(frames):
        pc = camm.make_push_constant(cam, h, frame_count=f, total_samples=total, num_samples=num_samples, max_depth=max_depth)
        pc.flags |= abi.PT_USE_OPTIX_DENOISER
        o.render_frame(fi, pc, accum, guide=guide)
        total += num_samples
    return accum, guide[..., :3].copy(), guide[..., 3].copy().view(np.uint32)


def test_oracle_guide_known_answers(box_scene, std_env, oracle_mod):
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    o.set_environment(std_env)
    accum, albedo, packed = render_with_guide(oracle_mod, o, box_scene.camera, 96, 64, 2, num_samples=2)
    # frame 1's last sample decides; .w of a 1-frame render tells which pixels are covered by all / none of the samples
    acc1 = np.zeros((64, 96, 4), np.float32)
    from vk_gltf_renderer_b200 import camera as camm
    fi = camm.make_frame_info(box_scene.camera, 96, 64)
    o.render_frame(fi, camm.make_push_constant(box_scene.camera, 64, frame_count=1, total_samples=0, num_samples=2, max_depth=4), acc1)
    miss, hit = acc1[..., 3] == 0.0, acc1[..., 3] == 1.0
    assert miss.sum() > 500 and hit.sum() > 500
    assert (albedo[miss] == 0).all() and (packed[miss] == 0x7FFF7FFF).all()        # (0, 0, 1): x = y = 0 on the upper octahedron
    m = box_scene.materials[0]
    base = np.float32(list(m.pbrBaseColorFactor)[:3]).astype(np.float16).astype(np.float32)
    assert np.array_equal(np.unique(albedo[hit].reshape(-1, 3), axis=0), base[None])  # untextured Box: the factor as binary16
    n = decode_unit_vec(packed[hit])
    assert np.allclose(np.linalg.norm(n, axis=-1), 1.0) and (n[:, 2] > 0).mean() > 0.99   # shading normals face the camera (-z forward)
    # the cube shows three faces at most: at most three distinct normals up to quantisation
    q = np.round(n * 50).astype(int)
    assert len(np.unique(q, axis=0)) <= 6
    # encoding round trip on random directions: 16 bits per axis
    rng = np.random.default_rng(0)
    v = rng.normal(size=(1000, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    d = 32767.0 / np.abs(v).sum(1)
    x, y = np.round(v[:, 0] * d).astype(np.int64), np.round(v[:, 1] * d).astype(np.int64)
    lo = v[:, 2] < 0
    mx, my = x >> 63, y >> 63
    tmp = 32767 + mx + my
    x2 = np.where(lo, (tmp - (y ^ my)) ^ mx, x)
    y2 = np.where(lo, (tmp - (x ^ mx)) ^ my, y)
    back = decode_unit_vec(((y2 + 32767) << 16 | (x2 + 32767)).astype(np.uint32))
    assert np.abs(back - v).max() < 2e-4
