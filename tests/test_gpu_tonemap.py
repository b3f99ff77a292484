"""-m gpu: tone mapping + 8-bit encode of the accumulation image through the C-ABI (b200pt_tonemap / b200pt_tonemap_image;
GltfRenderer::tonemap, src/renderer.cpp:992-1054) against oracle/tonemap.py on the image the device rendered."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _close(got, ref, frac=5e-3):
    d = np.abs(got.astype(int) - ref.astype(int))
    return d.max() <= 1 and (d > 0).mean() < frac


def test_tonemap_of_a_rendered_frame_matches_the_oracle(std_env):
    import torch
    from oracle import tonemap as T
    from vk_gltf_renderer_b200 import synth
    from vk_gltf_renderer_b200.renderer import B200PTError, PathTracer, Resources, render_headless
    scn = synth.synth_lit()
    res = Resources(scene=scn, hdr_rgb=std_env, camera=scn.camera, size=(192, 128))
    pt, img = render_headless(res, 6, ptMaxDepth=5)
    assert np.isfinite(img).all() and img[..., :3].max() > 1.0     # the bright lights need the operator
    for method in range(6):
        tm = pt.make_tonemapper(method=method, exposure=0.9, brightness=1.05, contrast=1.1, saturation=0.9, vignette=0.2)
        got, ex = pt.tonemap(tm)
        ref, _ = T.tonemap(img, method=method, exposure=0.9, brightness=1.05, contrast=1.1, saturation=0.9, vignette=0.2)
        assert abs(ex - 0.9) < 1e-7 and _close(got, ref), method
    # auto exposure (the reference's default): same histogram, same factor, same picture
    tm = pt.make_tonemapper(method=0, autoExposure=1)
    got, ex = pt.tonemap(tm)
    ref, ex_ref = T.tonemap(img, method=0, auto=1)
    assert abs(ex / float(ex_ref) - 1.0) < 1e-5 and _close(got, ref)
    assert 0.25 < got[..., :3].mean() / 255.0 < 0.75                 # exposed to the middle
    # inactive tonemapper: clamped linear colour
    got, _ = pt.tonemap(pt.make_tonemapper(isActive=0))
    assert np.array_equal(got, T.tonemap(img, is_active=0)[0])
    # any device image (e.g. the gathered multi-GPU frame): same kernel
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.empty((128, 192, 4), dtype=torch.uint8, device="cuda")
    tm = pt.make_tonemapper(method=3, vignette=0.3)
    pt.tonemap_image(tm, d_in.data_ptr(), 192, 128, d_out.data_ptr())
    assert _close(d_out.cpu().numpy(), T.tonemap(img, method=3, vignette=0.3)[0])
    # a row tile uses the full frame's coordinates for the vignette
    res2 = Resources(scene=scn, hdr_rgb=std_env, camera=scn.camera, size=(192, 128), tile=(32, 64))
    pt2, img2 = render_headless(res2, 2, ptMaxDepth=3)
    got2, _ = pt2.tonemap(pt2.make_tonemapper(method=5, vignette=0.4))
    assert _close(got2, T.tonemap(img2, method=5, vignette=0.4, y0=32, full_height=128)[0])
    with pytest.raises(B200PTError):
        pt.tonemap(pt.make_tonemapper(method=9))
    pt.onDetach(res)
    pt2.onDetach(res2)


def test_headless_driver_writes_the_tonemapped_image(tmp_path):
    """`--output file` of the reference application (src/renderer.cpp:171, saveHeadlessOutputImage :557-573): the C++ headless
    driver tonemaps its accumulation image with Resources::tonemapperData's defaults (filmic, auto-exposure ON,
    src/resources.hpp:212) through b200pt_tonemap and writes the 8-bit pixels (PPM: no jpg encoder here); they equal the
    oracle's tonemap of the RGBA32F image the same run wrote."""
    import os
    import subprocess
    from oracle import tonemap as T
    from vk_gltf_renderer_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "b200pt_headless")
    raw, ppm = str(tmp_path / "box.raw"), str(tmp_path / "box.jpg")
    cmd = [exe, "--headless", "--size", "96", "64", "--frames", "4", "--scenefile", os.path.join(root, "tests", "assets", "Box.glb"),
           "--hdrfile", os.path.join(root, "tests", "assets", "std_env.hdr"), "--ptMaxDepth", "4", "--outRaw", raw, "--output", ppm]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "HEADLESS_OUTPUT" in out.stdout, out.stdout
    img = np.fromfile(raw, np.float32).reshape(64, 96, 4)
    with open(ppm, "rb") as f:
        assert f.readline() == b"P6\n" and f.readline() == b"96 64\n" and f.readline() == b"255\n"
        got = np.frombuffer(f.read(), np.uint8).reshape(64, 96, 3)
    ref, ex = T.tonemap(img, method=0, auto=1)
    line = [l for l in out.stdout.splitlines() if l.startswith("HEADLESS_OUTPUT")][0]
    assert abs(float(line.split(" exposure=")[1]) / float(ex) - 1.0) < 1e-4
    assert _close(got, ref[..., :3])
