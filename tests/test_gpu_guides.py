"""-m gpu: the denoiser guide image (b200pt_set_guide_outputs / b200pt_read_guide; OutputImage::eOptixAlbedoNormal,
shaders/gltf_pathtrace.slang:240-263, 653-670) against the oracle on the textured material zoo: guide albedo (binary16 values of
the evaluated base colour) and the compressed camera-space shading normal of the newest sample's first hit."""
import numpy as np
import pytest

from conftest import rel_rmse
from test_guides import decode_unit_vec, render_with_guide

pytestmark = pytest.mark.gpu


def test_guide_image_matches_the_oracle(std_env, oracle_mod):
    from vk_gltf_renderer_b200 import synth
    from vk_gltf_renderer_b200.renderer import B200PTError, PathTracer, Resources
    scn = synth.synth_material_zoo()
    o = oracle_mod.Oracle()
    o.set_scene(scn)
    o.set_environment(std_env)
    W, H = 192, 128
    res = Resources(scene=scn, hdr_rgb=std_env, camera=scn.camera, size=(W, H))
    pt = PathTracer(0)
    pt.ptMaxDepth = 4
    pt.ptSamples = 2
    pt.onAttach(res)
    pt.ptUseOptixDenoiser = True
    res.frameCount = 0
    with pytest.raises(B200PTError):
        pt.onRender(None, res)                      # the flag without the guide image
    pt.set_guide_outputs(True)
    for batch in (1, 3):
        pt.set_frame_batch(batch)
        for f in range(3):
            res.frameCount = f
            pt.onRender(None, res)
        img = pt.read_accum()
        albedo, packed = pt.read_guide()
        ref_img, ref_albedo, ref_packed = render_with_guide(oracle_mod, o, scn.camera, W, H, 3, num_samples=2, max_depth=4)
        assert rel_rmse(img, ref_img) <= 1e-3       # the guides change nothing in the image
        # binary16 values of fp32 results that agree to an ulp or so: equal except at a rounding boundary
        same = (albedo == ref_albedo).all(-1)
        assert same.mean() > 0.995 and np.abs(albedo - ref_albedo).max() < 2e-3, (batch, same.mean())
        assert (albedo.astype(np.float16).astype(np.float32) == albedo).all()
        n, n_ref = decode_unit_vec(packed), decode_unit_vec(ref_packed)
        assert (packed == ref_packed).mean() > 0.99 and np.abs(n - n_ref).max() < 2e-3, (batch, (packed == ref_packed).mean())
        hitmask = (ref_albedo != 0).any(-1)
        assert hitmask.mean() > 0.4 and (packed[~hitmask & (ref_packed == 0x7FFF7FFF)] == 0x7FFF7FFF).all()
    pt.set_guide_outputs(False)
    pt.set_frame_batch(1)
    res.frameCount = 0
    pt.onRender(None, res)                          # back to the plain variant
    with pytest.raises(B200PTError):
        pt.read_guide()
    pt.onDetach(res)
