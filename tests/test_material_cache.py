"""The reference's own pins for the loader side of the path ("adjacent pins only"): the expectations of
/root/reference/tests/test_material_cache.cpp:24-176 (alpha modes, factors, texture-slot sentinel, topology-change detection)
restated against scene.MaterialCache, and the layout anchors src/gltf_material_cache.cpp:46-56 against the ABI struct."""
import ctypes as C

import pytest

from vk_gltf_renderer_b200 import abi, scene


def test_build_from_empty_materials():  # test_material_cache.cpp:24-32
    cache = scene.MaterialCache()
    cache.buildFromMaterials([])
    assert cache.getShadeMaterials() == []
    assert len(cache.getTextureInfos()) == 1  # sentinel entry at index 0


def test_build_from_single_opaque_material():  # :34-57
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{"alphaMode": "OPAQUE", "doubleSided": False,
                               "pbrMetallicRoughness": {"baseColorFactor": [1.0, 0.0, 0.0, 1.0], "metallicFactor": 0.5, "roughnessFactor": 0.8}}])
    assert len(cache.getShadeMaterials()) == 1
    m = cache.getShadeMaterials()[0]
    assert m.alphaMode == 0 and m.doubleSided == 0
    assert m.pbrBaseColorFactor[0] == pytest.approx(1.0) and m.pbrBaseColorFactor[1] == pytest.approx(0.0)
    assert m.pbrMetallicFactor == pytest.approx(0.5) and m.pbrRoughnessFactor == pytest.approx(0.8)


def test_mask_and_blend_alpha_modes():  # :59-77
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{"alphaMode": "MASK", "alphaCutoff": 0.3}, {"alphaMode": "BLEND"}])
    ms = cache.getShadeMaterials()
    assert len(ms) == 2
    assert ms[0].alphaMode == 1 and ms[0].alphaCutoff == pytest.approx(0.3)
    assert ms[1].alphaMode == 2


def test_base_color_texture_gets_a_texture_info():  # :79-95
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0, "texCoord": 0}}}])
    assert len(cache.getShadeMaterials()) == 1
    assert cache.getShadeMaterials()[0].pbrBaseColorTexture > 0  # > 0: a texture info was added
    assert len(cache.getTextureInfos()) >= 2                     # sentinel + the real one
    assert cache.getTextureInfos()[0].index == -1


def test_update_material_in_place():  # :97-115
    cache = scene.MaterialCache()
    mat = {"alphaMode": "OPAQUE", "pbrMetallicRoughness": {"roughnessFactor": 0.5, "baseColorTexture": {"index": 0}}}
    cache.buildFromMaterials([mat])
    slot = cache.getShadeMaterials()[0].pbrBaseColorTexture
    mat = {"alphaMode": "OPAQUE", "pbrMetallicRoughness": {"roughnessFactor": 0.9, "baseColorTexture": {"index": 0}}}
    r = cache.updateMaterial(0, mat)
    assert not r.topologyChanged
    assert cache.getShadeMaterials()[0].pbrRoughnessFactor == pytest.approx(0.9)
    assert cache.getShadeMaterials()[0].pbrBaseColorTexture == slot


def test_update_detects_topology_change_when_texture_added_or_removed():  # :117-151
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{"alphaMode": "OPAQUE"}])
    assert cache.updateMaterial(0, {"alphaMode": "OPAQUE", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}).topologyChanged
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}])
    assert cache.updateMaterial(0, {"pbrMetallicRoughness": {"baseColorTexture": {"index": -1}}}).topologyChanged
    assert cache.getShadeMaterials()[0].pbrBaseColorTexture == 0 and len(cache.getTextureInfos()) == 1


def test_update_out_of_range_returns_empty():  # :153-164
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{}])
    r = cache.updateMaterial(5, {})
    assert not r.topologyChanged and not r.hasAny()


def test_clear_resets_all():  # :166-176
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{}])
    cache.clear()
    assert cache.getShadeMaterials() == [] and cache.getTextureInfos() == []


def test_texcoord_is_clamped_to_the_two_sets_the_shaders_know():
    cache = scene.MaterialCache()
    cache.buildFromMaterials([{"emissiveTexture": {"index": 2, "texCoord": 5}}])
    ti = cache.getTextureInfos()[cache.getShadeMaterials()[0].emissiveTexture]
    assert ti.index == 2 and ti.texCoord == 1


def test_layout_anchors_of_the_shade_material():  # src/gltf_material_cache.cpp:46-56
    M = abi.ShadeMaterial
    assert C.sizeof(M) % 8 == 0 and C.alignment(M) >= 8
    assert M.pbrBaseColorFactor.offset == 0
    assert M.pbrRoughnessFactor.offset == 32
    assert M.alphaMode.offset == 40
    assert M.occlusionStrength.offset == 48
    assert M.doubleSided.offset == 52


def test_loader_uses_the_same_conversion(box_scene):
    """load_gltf goes through populate_shade_material: Box.glb's single material (red-ish, metallic 0) arrives unchanged."""
    m = box_scene.materials[0]
    assert m.alphaMode == 0 and m.pbrMetallicFactor == pytest.approx(0.0)
    assert tuple(round(x, 3) for x in m.pbrBaseColorFactor) == (0.8, 0.0, 0.0, 1.0)
