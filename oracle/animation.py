"""CPU restatement (numpy, fp32, operation by operation) of the reference's animation compute shaders -- TEST INFRASTRUCTURE,
never imported by the product.

  morph(...)  shaders/morph.comp.slang:29-70
  skin(...)   shaders/skinning.comp.slang:27-70
  apply(...)  the order SceneAnimationVk::cmdUpdateAnimation dispatches them in (src/gltf_scene_animation_vk.cpp:497-582):
              all morph tasks, then all skin tasks; a primitive that is both is skinned from its morphed arrays (:545-556)

Arithmetic order (the reference leaves it to the SPIR-V compiler; pinned identically in csrc/animate.cuh): every product and sum
is a separate fp32 operation, a matrix row is ((m0*x + m1*y) + m2*z) [+ m3], normalize(v) = v / sqrt((x*x + y*y) + z*z).
Matrices arrive in glm byte order: M[j, c, r] = column c, row r (mul(v, M) in the shaders is M_glm * v).
Parity with the reference: the shader sources are in the tree and are followed statement by statement; there is no
reference-held golden vector for them (no test of the reference runs a shader).
"""
import numpy as np

F = np.float32


def _normalize(v):
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        l = np.sqrt((x * x + y * y) + z * z)
        return (v / l[:, None]).astype(F)


def morph(base_pos, base_nrm, base_tan, d_pos, d_nrm, d_tan, weights):
    """-> (positions, normals or None, tangents or None)"""
    pos = np.array(base_pos, F)
    has_n, has_t = base_nrm is not None, base_tan is not None
    nrm = np.array(base_nrm, F) if has_n else None
    tan = np.array(base_tan, F)[:, :3].copy() if has_t else None
    for t, w in enumerate(np.asarray(weights, F)):
        if w == 0.0:
            continue
        pos = pos + w * np.asarray(d_pos[t], F)
        if has_n and d_nrm is not None:
            nrm = nrm + w * np.asarray(d_nrm[t], F)
        if has_t and d_tan is not None:
            tan = tan + w * np.asarray(d_tan[t], F)
    out_n = _normalize(nrm) if has_n else None
    out_t = np.concatenate([_normalize(tan), np.asarray(base_tan, F)[:, 3:4]], 1) if has_t else None
    return pos.astype(F), out_n, out_t


def skin(base_pos, base_nrm, base_tan, weights, joints, joint_mats, normal_mats):
    p = np.asarray(base_pos, F)
    has_n, has_t = base_nrm is not None, base_tan is not None
    n = np.asarray(base_nrm, F) if has_n else None
    t = np.asarray(base_tan, F) if has_t else None
    jm, nm = np.asarray(joint_mats, F), np.asarray(normal_mats, F)
    nj = jm.shape[0]
    w, j = np.asarray(weights, F), np.asarray(joints, np.int32)
    sp = np.zeros_like(p)
    sn = np.zeros_like(p)
    st = np.zeros_like(p)
    for i in range(4):
        jw, ji = w[:, i], j[:, i]
        ok = (jw > 0.0) & (ji >= 0) & (ji < nj)
        M = jm[np.clip(ji, 0, nj - 1)]  # [V, c, r]
        N = nm[np.clip(ji, 0, nj - 1)]

        def rows(A, v, k, translate):
            out = []
            for r in range(3):
                acc = (A[:, 0, r] * v[:, 0] + A[:, 1, r] * v[:, 1]) + A[:, 2, r] * v[:, 2]
                out.append(acc + A[:, 3, r] if translate else acc)
            return np.stack(out, 1).astype(F)
        q = rows(M, p, 4, True)
        sp = np.where(ok[:, None], sp + jw[:, None] * q, sp).astype(F)
        if has_n:
            r_ = rows(N, n, 3, False)
            sn = np.where(ok[:, None], sn + jw[:, None] * r_, sn).astype(F)
        if has_t:
            r_ = rows(M, t, 3, False)
            st = np.where(ok[:, None], st + jw[:, None] * r_, st).astype(F)
    out_n = _normalize(sn) if has_n else None
    out_t = np.concatenate([_normalize(st), t[:, 3:4]], 1).astype(F) if has_t else None
    return sp, out_n, out_t


def apply(scene, morph_tasks, skin_tasks, morph_weights, joint_mats, normal_mats):
    """Updates scene.render_prims[*]['positions' / 'normals' / 'tangents'] in place the way one cmdUpdateAnimation does."""
    morphed = set()
    for t, w in zip(morph_tasks, morph_weights):
        prim = scene.render_prims[t.render_prim]
        if len(w) == 0:
            continue
        bn = t.base_normals if prim.get("normals") is not None else None
        bt = t.base_tangents if prim.get("tangents") is not None else None
        pos, nrm, tan = morph(t.base_positions, bn, bt, t.position_deltas, t.normal_deltas if bn is not None else None,
                              t.tangent_deltas if bt is not None else None, w)
        prim["positions"] = pos
        if nrm is not None:
            prim["normals"] = nrm
        if tan is not None:
            prim["tangents"] = tan
        morphed.add(t.render_prim)
    for t, jm, nm in zip(skin_tasks, joint_mats, normal_mats):
        prim = scene.render_prims[t.render_prim]
        has_n, has_t = prim.get("normals") is not None, prim.get("tangents") is not None
        if t.render_prim in morphed:
            bp, bn, bt = prim["positions"], prim["normals"] if has_n else None, prim["tangents"] if has_t else None
        else:
            bp, bn, bt = t.base_positions, t.base_normals if has_n else None, t.base_tangents if has_t else None
        pos, nrm, tan = skin(bp, bn, bt, t.weights, t.joints, jm, nm)
        prim["positions"] = pos
        if nrm is not None:
            prim["normals"] = nrm
        if tan is not None:
            prim["tangents"] = tan
