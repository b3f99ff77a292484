"""Tone mapping + 8-bit encode without a GPU: known answers of the numpy restatement (oracle/tonemap.py) and the DEVICE source
(csrc/tonemap.cuh, the bodies of k_tonemap / k_tm_histogram) compiled for the host and compared with it.  Parity with the
reference's external nvshaders tonemapper is unpinned; these are the operators' published properties."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vk_gltf_renderer_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
sys.path.insert(0, ROOT)


def _test_image(rows=48, width=64, seed=3):
    """radiance over 14 stops, some black, some huge, a few negative / NaN-free oddities, alpha 0 / fractional / 1"""
    rng = np.random.default_rng(seed)
    img = (2.0 ** rng.uniform(-10, 4, size=(rows, width, 1)) * rng.uniform(0.2, 1.0, size=(rows, width, 3))).astype(np.float32)
    img[:4] = 0.0
    img[4:6] *= 1e4
    img[6, :8] = -0.25
    a = rng.choice([0.0, 0.5, 1.0, 0.9999], size=(rows, width, 1)).astype(np.float32)
    return np.concatenate([img, a], -1)


def test_operator_known_answers():
    from oracle import tonemap as T
    g = np.linspace(0, 1, 33, dtype=np.float32)
    gray = np.stack([g, g, g], -1)
    # sRGB anchors: 0 -> 0, 1 -> 1, the linear segment below 0.0031308, 0.18 -> 0.4614 (mid grey)
    assert T.srgb(np.float32([0.0, 1.0]))[0] == 0.0 and abs(float(T.srgb(np.float32([1.0]))[0]) - 1.0) < 1e-6
    assert np.allclose(T.srgb(np.float32([0.001])), 0.01292, atol=1e-7) and abs(float(T.srgb(np.float32([0.18]))[0]) - 0.46135613) < 2e-6
    x = np.float32(2.0) ** np.linspace(-8, 6, 57, dtype=np.float32)
    ramp = np.stack([x, x, x], -1)
    for method in range(6):
        out = T.operator(method, ramp)
        assert np.isfinite(out).all()
        assert (np.diff(out[:, 0]) >= -1e-6).all(), method             # monotone on a grey ramp
        assert np.allclose(out[:, 0], out[:, 1], atol=1e-5) and np.allclose(out[:, 0], out[:, 2], atol=1e-5) or method == 4  # grey stays grey
        assert out[0, 0] < 0.12 and out[-1, 0] > 0.75, (method, out[0, 0], out[-1, 0])
    # filmic: Hejl / Burgess-Dawson at 0.18: t = 0.176, t (6.2 t + 0.5) / (t (6.2 t + 1.7) + 0.06) = 0.28005 / 0.55125 (display-encoded)
    assert abs(float(T.operator(0, np.float32([[0.18, 0.18, 0.18]]))[0, 0]) - 0.28005 / 0.55125) < 1e-4
    # Uncharted 2: the white point 11.2 / exposure bias 2 maps linear 5.6 to display 1
    assert abs(float(T.operator(1, np.float32([[5.6, 5.6, 5.6]]))[0, 0]) - 1.0) < 1e-5
    # Khronos PBR neutral: below the compression start colours pass through after the toe offset (published property):
    # base colour 0.5 grey -> sRGB(0.5 - 0.04)
    assert np.allclose(T.operator(5, np.float32([[0.5, 0.5, 0.5]])), T.srgb(np.float32([0.46])), atol=1e-6)
    # ... and the output never exceeds 1
    assert T.operator(5, ramp * 50).max() <= 1.0 + 1e-6
    # clip == sRGB encode
    assert np.array_equal(T.operator(2, gray), T.srgb(gray))


def test_post_controls_and_encode():
    from oracle import tonemap as T
    img = _test_image()
    base, ex = T.tonemap(img, method=2)
    assert ex == 1.0 and base.dtype == np.uint8 and base.shape == img.shape
    assert (base[:4, :, :3] == 0).all()                      # black stays black
    assert (base[6, :8, :3] == 0).all()                      # negative radiance clamps to 0
    assert set(np.unique(base[..., 3])) <= {0, 128, 255}     # alpha: UNORM8 of the coverage flag mean (0.5 -> 128, 0.9999 -> 255)
    assert np.array_equal(T.tonemap(img, method=2, is_active=0)[0][..., :3],
                          (np.clip(img[..., :3], 0, 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8))
    # saturation 0 gives grey pixels; vignette darkens the corners, not the centre
    grey = T.tonemap(img, method=0, saturation=0.0)[0].astype(int)
    assert (np.abs(grey[..., 0] - grey[..., 1]) <= 1).all() and (np.abs(grey[..., 0] - grey[..., 2]) <= 1).all()
    flat = np.full((32, 32, 4), 0.5, np.float32)
    vg = T.tonemap(flat, method=2, vignette=0.5)[0]
    assert vg[0, 0, 0] < vg[16, 16, 0] and abs(int(vg[16, 16, 0]) - int(T.tonemap(flat, method=2)[0][16, 16, 0])) <= 1
    # a row tile with y0 / full_height reproduces the rows of the full image
    full = T.tonemap(img, method=3, vignette=0.3)[0]
    tile = T.tonemap(img[16:32], method=3, vignette=0.3, y0=16, full_height=img.shape[0])[0]
    assert np.array_equal(tile, full[16:32])
    # auto exposure: a constant image of luminance L is exposed to 0.18 (to the histogram's bin resolution of 1/8 stop)
    const = np.full((16, 16, 4), 0.03, np.float32)
    ex, hist = T.auto_exposure(const, 1.0)
    assert hist.sum() == 256 and abs(np.log2(float(ex) * 0.03 / 0.18)) <= 0.0626
    assert T.auto_exposure(np.zeros((4, 4, 4), np.float32), 2.0)[0] == 2.0   # all black: the base exposure


def test_device_tonemap_source_matches_the_oracle(tmp_path):
    """csrc/tonemap.cuh compiled for the host (tools/host_tonemap_check.cpp): every operator with non-trivial post controls on a
    14-stop test image; 8-bit results equal the oracle's except where libm and numpy round a transcendental differently (<= 1 LSB,
    < 0.5 % of the values), the histogram bins are identical."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from oracle import tonemap as T
    from vk_gltf_renderer_b200 import abi
    exe = str(tmp_path / "host_tonemap_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + CUDA_INC, "-o", exe, os.path.join(CSRC, "tools", "host_tonemap_check.cpp")])
    img = _test_image()
    rows, width = img.shape[:2]
    for method in range(6):
        for active in (1, 0):
            tm = abi.Tonemapper(method, active, 1.3, 1.1, 1.2, 0.8, 0.25, 0)
            with open(tmp_path / "in.bin", "wb") as f:
                f.write(np.array([width, rows // 2, 8, rows], np.uint32).tobytes())
                f.write(bytes(tm))
                f.write(np.float32(tm.exposure).tobytes())
                f.write(img[8:8 + rows // 2].tobytes())
            subprocess.check_call([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
            raw = np.fromfile(str(tmp_path / "out.bin"), np.uint8)
            got = raw[:rows // 2 * width * 4].reshape(rows // 2, width, 4)
            hist = raw[rows // 2 * width * 4:].view(np.uint32)
            ref, _ = T.tonemap(img[8:8 + rows // 2], method=method, is_active=active, exposure=1.3, brightness=1.1, contrast=1.2, saturation=0.8,
                               vignette=0.25, y0=8, full_height=rows)
            d = np.abs(got.astype(int) - ref.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 5e-3, (method, active, d.max(), (d > 0).mean())
            _, href = T.auto_exposure(img[8:8 + rows // 2], 1.0)
            h0 = hist.astype(np.float64).copy()
            h0[0] = 0
            assert np.abs(h0 - href).sum() <= 2, method     # a luminance on a bin edge may land on either side
