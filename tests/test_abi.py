"""CPU: the C-ABI library loads, exports every symbol include/b200pt.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200pt_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_match_python_binding():
    from vk_gltf_renderer_b200 import _lib
    assert declared_symbols() == sorted(_lib.EXPORTS)


@pytest.mark.parametrize("variant", [False, True])
def test_library_exports_every_declared_symbol(variant):
    from vk_gltf_renderer_b200 import _lib
    L = _lib.lib(count_traversal=variant)
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert L.b200pt_abi_version() == 5


def test_struct_layouts_match_reference_sizes():
    from vk_gltf_renderer_b200 import abi
    # sizes / anchors from shaders/gltf_scene_io.h.slang, shaders/shaderio.h, src/gltf_material_cache.cpp:46-56
    assert C.sizeof(abi.RenderNode) == 136 and C.sizeof(abi.ShadeMaterial) == 288
    assert C.sizeof(abi.TextureInfo) == 32 and C.sizeof(abi.Light) == 64
    assert C.sizeof(abi.FrameInfo) == 396 and C.sizeof(abi.PushConstant) == 48
    m = abi.ShadeMaterial
    assert (m.pbrBaseColorFactor.offset, m.pbrRoughnessFactor.offset, m.alphaMode.offset, m.occlusionStrength.offset,
            m.doubleSided.offset) == (0, 32, 40, 48, 52)
    assert m.pbrBaseColorTexture.offset == 232 and m.retroreflectionTexture.offset == 274


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device b200pt_create must fail; the Python host raises instead of rendering on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vk_gltf_renderer_b200.renderer import B200PTError, PathTracer, Resources
    with pytest.raises(B200PTError):
        PathTracer(0).onAttach(Resources(size=(8, 8)))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, link or load it."""
    pkg = os.path.join(ROOT, "vk_gltf_renderer_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|liboracle|#include\s+\"[^\"]*oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), os.path.join(dirpath, f)
