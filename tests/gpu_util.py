"""Helpers shared by the -m gpu tests: device buffers through torch (plumbing only)."""
import numpy as np


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def primary_rays(cam, w, h, jitter=0.5):
    """Pinhole camera rays (o,tmin,d,tmax) for ray-level parity tests."""
    from vk_gltf_renderer_b200 import camera as cm
    import math
    view = cm.look_at(cam.eye, cam.center, cam.up)
    vinv = np.linalg.inv(view)
    t = math.tan(cam.yfov / 2)
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    cx = ((xs + jitter) / w * 2 - 1) * t * (w / h)
    cy = -((ys + jitter) / h * 2 - 1) * t
    d = np.stack([cx, cy, -np.ones_like(cx)], -1).reshape(-1, 3)
    d = d @ vinv[:3, :3].T
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((w * h, 8), np.float32)
    rays[:, 0:3] = vinv[:3, 3]
    rays[:, 3] = 0.0
    rays[:, 4:7] = d
    rays[:, 7] = 1e32
    return rays


def random_rays(n, lo, hi, seed=1234):
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    c, r = (lo + hi) / 2, np.linalg.norm(hi - lo) / 2
    o = c + rng.normal(size=(n, 3)) * r
    tgt = c + (rng.random((n, 3)) - 0.5) * (hi - lo)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 4:7] = d
    rays[:, 7] = 1e32
    return rays
