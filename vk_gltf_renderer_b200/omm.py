"""Opacity micromaps: the EXT_mesh_opacity_micromap arrays of a scene, and a baker for assets that carry none.

Reference: src/gltf_scene_omm.{hpp,cpp} consumes PRE-BAKED micromaps from the asset (EXT_mesh_opacity_micromap) and hands them to
VK_EXT_opacity_micromap; there is no baker in the reference tree (assets are baked offline, e.g. with NVIDIA's OMM SDK).  The
software walk of this package consumes the same arrays (csrc/omm.cuh).  `bake_opacity_micromaps` plays the offline tool's part for
the synthetic stand-in scenes: it classifies every micro-triangle of every alpha-MASK triangle from the base-colour texture,
CONSERVATIVELY -- a state is OPAQUE / TRANSPARENT only when every texel the any-hit evaluation could touch for a hit inside the
micro-triangle (level-0 texels of its uv bounding box, one texel of margin for the bilinear footprint and rounding) is on the same
side of the cutoff, with a relative safety band of 1e-4; everything else is UNKNOWN and still runs getOpacity.  So the image with
the baked micromaps differs from the one without only in which rand() calls are spent on any-hit candidates (exactly as with the
hardware feature: raytracer_interface.h.slang:93-100).

Micro-triangle order: VK_EXT_opacity_micromap's bary2index (restated from the specification; see csrc/omm.cuh for what pins it).
"""
import numpy as np

from . import abi


def bary2index(u, v, level):
    """micro-triangle index of barycentrics (u, v) (weights of the 2nd / 3rd vertex) at a subdivision level; vectorised"""
    u = np.clip(np.asarray(u, np.float32), 0, 1)
    v = np.clip(np.asarray(v, np.float32), 0, 1)
    n = np.uint32(1 << level)
    fu, fv = u * np.float32(n), v * np.float32(n)
    iu, iv = fu.astype(np.uint32), fv.astype(np.uint32)
    uf, vf = fu - iu.astype(np.float32), fv - iv.astype(np.float32)
    iu, iv = np.minimum(iu, n - np.uint32(1)), np.minimum(iv, n - np.uint32(1))
    iuv = iu.astype(np.int64) + iv.astype(np.int64)
    iu = np.where(iuv >= int(n), iu.astype(np.int64) - (iuv - int(n) + 1), iu.astype(np.int64)).astype(np.uint32)
    iw = (~(iu + iv)).astype(np.uint32)
    dec = ((uf + vf) >= np.float32(1.0)) & (iuv < int(n) - 1)
    iw = np.where(dec, iw - np.uint32(1), iw).astype(np.uint32)
    mask = np.uint32(int(n) - 1)
    b0 = (~(iu ^ iw)) & mask
    t = (iu ^ iv) & b0
    f = t.copy()
    for s in (1, 2, 4, 8):
        f = f ^ (f >> np.uint32(s))
    b1 = ((f ^ iu) & ~b0) | t

    def spread(x):
        x = x.astype(np.uint32)
        x = (x | (x << np.uint32(8))) & np.uint32(0x00ff00ff)
        x = (x | (x << np.uint32(4))) & np.uint32(0x0f0f0f0f)
        x = (x | (x << np.uint32(2))) & np.uint32(0x33333333)
        x = (x | (x << np.uint32(1))) & np.uint32(0x55555555)
        return x
    return (spread(b0) | (spread(b1) << np.uint32(1))).astype(np.uint32)


_corner_cache = {}


def micro_triangle_corners(level):
    """(4^level, 3, 2) barycentric (u, v) corners of every micro-triangle, row = its bary2index"""
    if level not in _corner_cache:
        n = 1 << level
        corners, cents = [], []
        for i in range(n):
            for j in range(n - i):
                corners.append([(i, j), (i + 1, j), (i, j + 1)])
                cents.append(((i + 1 / 3) / n, (j + 1 / 3) / n))
                if i + j < n - 1:
                    corners.append([(i + 1, j), (i + 1, j + 1), (i, j + 1)])
                    cents.append(((i + 2 / 3) / n, (j + 2 / 3) / n))
        corners = np.asarray(corners, np.float64) / n
        cents = np.asarray(cents, np.float32)
        idx = bary2index(cents[:, 0], cents[:, 1], level)
        out = np.empty_like(corners)
        out[idx] = corners
        assert len(np.unique(idx)) == 4 ** level
        _corner_cache[level] = out
    return _corner_cache[level]


_cell_cache = {}


def micro_triangle_cells(level):
    """(i, j, upper) of every micro-triangle in bary2index order: the grid cell (u in [i, i+1) / n, v in [j, j+1) / n) it lies in and
    whether it is the cell's upper triangle (corners (i+1,j) (i+1,j+1) (i,j+1)) or the lower one ((i,j) (i+1,j) (i,j+1))"""
    if level not in _cell_cache:
        n = 1 << level
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        keep = (ii + jj) < n
        li, lj = ii[keep], jj[keep]
        ku = (ii + jj) < n - 1
        ui, uj = ii[ku], jj[ku]
        i = np.concatenate([li, ui])
        j = np.concatenate([lj, uj])
        up = np.concatenate([np.zeros(len(li), bool), np.ones(len(ui), bool)])
        cu = np.where(up, (i + 2 / 3) / n, (i + 1 / 3) / n).astype(np.float32)
        cv = np.where(up, (j + 2 / 3) / n, (j + 1 / 3) / n).astype(np.float32)
        idx = bary2index(cu, cv, level)
        order = np.argsort(idx)
        assert np.array_equal(idx[order], np.arange(4 ** level, dtype=np.uint32))
        _cell_cache[level] = (i[order].astype(np.float64), j[order].astype(np.float64), up[order])
    return _cell_cache[level]


def pack_states(states, fmt=abi.OMM_FORMAT_4_STATE):
    """(T, 4^level) uint8 states -> (T, bytes) packed like VkMicromap data (micro-triangle i in bits [i*b, i*b+b) of the stream)"""
    states = np.asarray(states, np.uint8)
    bits = 2 if fmt == abi.OMM_FORMAT_4_STATE else 1
    per = 8 // bits
    T, M = states.shape
    if T == 0:
        return np.zeros((0, (M * bits + 7) // 8), np.uint8)
    pad = (-M) % per
    if pad:
        states = np.concatenate([states, np.zeros((T, pad), np.uint8)], 1)
    s = states.reshape(T, -1, per).astype(np.uint32)
    shifts = (np.arange(per, dtype=np.uint32) * bits)[None, None, :]
    return (s << shifts).sum(-1).astype(np.uint8)


def _sat(mask):
    s = np.zeros((mask.shape[0] + 1, mask.shape[1] + 1), np.int64)
    s[1:, 1:] = mask.astype(np.int64).cumsum(0).cumsum(1)
    return s


def _rect(s, x0, y0, x1, y1):
    """sum over texels [x0, x1] x [y0, y1] (inclusive) of a summed-area table"""
    return s[y1 + 1, x1 + 1] - s[y0, x1 + 1] - s[y1 + 1, x0] + s[y0, x0]


def bake_opacity_micromaps(scene, level=4, fmt=abi.OMM_FORMAT_4_STATE, band=1e-4, refine=0, chunk=4096):
    """Fill scene.micromaps / scene.prim_omms for every primitive whose nodes all use one alpha-MASK material with a base-colour
    texture (no texture transform, no vertex-colour alpha, the triangle's uv inside one texture tile).  The texel boxes are taken
    `refine` levels finer than the micromap (a micro-triangle is known when all its 4^refine sub-triangles agree): tighter than one
    box around the whole micro-triangle.  Returns statistics."""
    scene.micromaps, scene.prim_omms = [], []
    fine = level + refine
    ci, cj, cup = micro_triangle_cells(fine)
    n = float(1 << fine)
    M, R = 4 ** level, 4 ** refine
    mats_of_prim = {}
    for rn in scene.render_nodes:
        mats_of_prim.setdefault(rn["renderPrimID"], set()).add(max(0, rn["materialID"]))
    sat_cache = {}
    data_chunks, tri_recs, offset = [], [], 0
    stats = dict(triangles=0, micro=0, opaque=0, transparent=0, unknown=0, fully_opaque=0, fully_transparent=0, level=level, refine=refine)
    bytes_per_tri = (M * (2 if fmt == abi.OMM_FORMAT_4_STATE else 1) + 7) // 8
    for pid, mids in sorted(mats_of_prim.items()):
        if len(mids) != 1:
            continue
        m = scene.materials[next(iter(mids))]
        if m.alphaMode != 1:
            continue
        spec_gloss = m.pbrModel == 1
        slot = m.pbrDiffuseTexture if spec_gloss else m.pbrBaseColorTexture
        factor = float(m.pbrDiffuseFactor[3] if spec_gloss else m.pbrBaseColorFactor[3])
        if slot <= 0 or factor <= 0.0:
            continue
        ti = scene.texture_infos[slot]
        if list(ti.uvTransform) != [1, 0, 0, 1, 0, 0] or ti.index < 0:
            continue
        prim = scene.render_prims[pid]
        uvs = prim["uv1"] if ti.texCoord else prim["uv0"]
        if uvs is None:
            continue
        if prim["colors"] is not None and np.any((prim["colors"] >> 24) != 255):
            continue
        tex = scene.textures[ti.index]
        a8 = tex["rgba8"][..., 3]
        H, W = a8.shape
        key = (ti.index, factor, float(m.alphaCutoff))
        if key not in sat_cache:
            a = a8.astype(np.float32) / np.float32(255.0) * np.float32(factor)
            sat_cache[key] = (_sat(a >= m.alphaCutoff * (1.0 + band)), _sat(a < m.alphaCutoff * (1.0 - band)))
        sat_op, sat_tr = sat_cache[key]
        tri_all = prim["indices"]
        states_all = []
        for c0 in range(0, len(tri_all), chunk):
            tri = tri_all[c0:c0 + chunk]
            T = len(tri)
            tuv = uvs[tri].astype(np.float64)                        # (T, 3, 2)
            base = np.floor(tuv.min(1))                              # the texture tile the triangle lies in (REPEAT)
            inside = np.all(tuv.max(1) <= base + 1.0, -1)
            if tex["wrapS"] != 10497 or tex["wrapT"] != 10497:
                inside &= np.all(base == 0.0, -1)
            size = np.array([W, H], np.float64)
            p0 = (tuv[:, 0] - base) * size                           # texel coordinates of the first vertex
            d1 = (tuv[:, 1] - tuv[:, 0]) * size / n                  # one grid step along u / v
            d2 = (tuv[:, 2] - tuv[:, 0]) * size / n
            zero = np.zeros_like(d1)
            lo_min = np.minimum(np.minimum(zero, d1), d2)            # corner offsets from the cell's (i, j) point: lower triangle
            lo_max = np.maximum(np.maximum(zero, d1), d2)
            up_min = np.minimum(np.minimum(d1, d1 + d2), d2)         # upper triangle
            up_max = np.maximum(np.maximum(d1, d1 + d2), d2)
            rect = []
            for ax, sz in ((0, W), (1, H)):
                pt = p0[:, ax, None] + ci[None, :] * d1[:, ax, None] + cj[None, :] * d2[:, ax, None]   # (T, 4^fine)
                mn = pt + np.where(cup[None, :], up_min[:, ax, None], lo_min[:, ax, None])
                mx = pt + np.where(cup[None, :], up_max[:, ax, None], lo_max[:, ax, None])
                rect.append((np.floor(mn).astype(np.int64) - 2, np.floor(mx).astype(np.int64) + 2, sz))
            (x0, x1, _), (y0, y1, _) = rect
            ok = inside[:, None] & (x0 >= 0) & (y0 >= 0) & (x1 < W) & (y1 < H)
            x0c, x1c, y0c, y1c = np.clip(x0, 0, W - 1), np.clip(x1, 0, W - 1), np.clip(y0, 0, H - 1), np.clip(y1, 0, H - 1)
            area = (x1c - x0c + 1) * (y1c - y0c + 1)
            n_op = _rect(sat_op, x0c, y0c, x1c, y1c)
            f_op = (ok & (n_op == area)).reshape(T, M, R)            # fine index >> (2 refine) = the micro-triangle (hierarchical order)
            f_tr = (ok & (_rect(sat_tr, x0c, y0c, x1c, y1c) == area)).reshape(T, M, R)
            mostly = (n_op * 2 >= area).reshape(T, M, R).mean(-1) >= 0.5
            states_all.append(np.where(f_op.all(-1), 1, np.where(f_tr.all(-1), 0, np.where(mostly, 3, 2))).astype(np.uint8))
        states = np.concatenate(states_all) if states_all else np.zeros((0, M), np.uint8)
        T = len(states)
        if fmt == abi.OMM_FORMAT_2_STATE:
            # two states cannot say "unknown": only fully known triangles may be linked
            known = np.all(states < 2, 1)
        else:
            known = np.ones(T, bool)
        full_op = np.all(states == 1, 1)
        full_tr = np.all(states == 0, 1)
        idx = np.full(T, abi.OMM_INDEX_FULLY_UNKNOWN_OPAQUE, np.int32)
        idx[full_op] = abi.OMM_INDEX_FULLY_OPAQUE
        idx[full_tr] = abi.OMM_INDEX_FULLY_TRANSPARENT
        need = known & ~full_op & ~full_tr
        nn = int(need.sum())
        idx[need] = len(tri_recs) + np.arange(nn, dtype=np.int32)
        packed = pack_states(states[need] & (3 if fmt == abi.OMM_FORMAT_4_STATE else 1), fmt)
        for k in range(nn):
            tri_recs.append((offset, level, fmt))
            offset += bytes_per_tri
        data_chunks.append(packed.reshape(-1))
        scene.prim_omms.append(dict(renderPrimID=pid, micromap=0, baseTriangle=0, indices=idx))
        stats["triangles"] += T
        stats["micro"] += T * M
        stats["opaque"] += int((states == 1).sum())
        stats["transparent"] += int((states == 0).sum())
        stats["unknown"] += int((states >= 2).sum())
        stats["fully_opaque"] += int(full_op.sum())
        stats["fully_transparent"] += int(full_tr.sum())
    if scene.prim_omms:
        data = np.concatenate(data_chunks) if data_chunks else np.zeros(0, np.uint8)
        scene.micromaps.append(dict(data=data, triangles=np.array(tri_recs, abi.MICROMAP_TRIANGLE_DTYPE)))
    scene._keep = []
    return stats
