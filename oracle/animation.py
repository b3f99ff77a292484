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


# ---- rigid part: shaders/world_matrix_propagate.comp.slang:27-42, shaders/update_render_instances.comp.slang:42-66 -------------
def mat4_mul(A, B):
    """glm product A * B on glm-ordered arrays [..., c, r]; C[c][r] = ((A[0][r] B[c][0] + A[1][r] B[c][1]) + A[2][r] B[c][2]) + A[3][r] B[c][3]"""
    A, B = np.asarray(A, F), np.asarray(B, F)
    C = np.empty(np.broadcast(A, B).shape, F)
    for c in range(4):
        for r in range(4):
            C[..., c, r] = ((A[..., 0, r] * B[..., c, 0] + A[..., 1, r] * B[..., c, 1]) + A[..., 2, r] * B[..., c, 2]) + A[..., 3, r] * B[..., c, 3]
    return C


def mat4_inverse(m):
    """glm's cofactor expansion (compute_inverse<4,4>) in fp32, same operation order as csrc/animate.cuh mat4Inverse"""
    m = np.asarray(m, F)
    M = lambda c, r: m[..., c, r]
    c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3); c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3); c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3)
    c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3); c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3); c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3)
    c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2); c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2); c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)
    c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3); c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3); c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3)
    c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2); c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2); c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2)
    c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1); c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1); c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)
    f0, f1, f2 = (c00, c00, c02, c03), (c04, c04, c06, c07), (c08, c08, c10, c11)
    f3, f4, f5 = (c12, c12, c14, c15), (c16, c16, c18, c19), (c20, c20, c22, c23)
    v0 = (M(1, 0), M(0, 0), M(0, 0), M(0, 0)); v1 = (M(1, 1), M(0, 1), M(0, 1), M(0, 1))
    v2 = (M(1, 2), M(0, 2), M(0, 2), M(0, 2)); v3 = (M(1, 3), M(0, 3), M(0, 3), M(0, 3))
    inv = np.empty(m.shape, F)
    for k in range(4):
        sa = F(-1.0) if (k & 1) else F(1.0)
        sb = -sa
        inv[..., 0, k] = ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]) * sa
        inv[..., 1, k] = ((v0[k] * f0[k] - v2[k] * f3[k]) + v3[k] * f4[k]) * sb
        inv[..., 2, k] = ((v0[k] * f1[k] - v1[k] * f3[k]) + v3[k] * f5[k]) * sa
        inv[..., 3, k] = ((v0[k] * f2[k] - v1[k] * f4[k]) + v2[k] * f5[k]) * sb
    d0, d1, d2, d3 = M(0, 0) * inv[..., 0, 0], M(0, 1) * inv[..., 1, 0], M(0, 2) * inv[..., 2, 0], M(0, 3) * inv[..., 3, 0]
    one_over_det = F(1.0) / ((d0 + d1) + (d2 + d3))
    return (inv * one_over_det[..., None, None]).astype(F)


def propagate(local, parents, order, offsets):
    """world[node] = world[parent] * local[node] level by level (identity for roots: the local matrix itself)"""
    local = np.asarray(local, F)
    world = np.zeros_like(local)
    for l in range(len(offsets) - 1):
        for n in order[offsets[l]:offsets[l + 1]]:
            p = parents[n]
            world[n] = local[n] if p < 0 else mat4_mul(world[p], local[n])
    return world


def render_nodes(world, mappings, inst_local=None):
    """(objectToWorld [N,4,4], worldToObject [N,4,4]) per render node: instLocal * world[nodeID] (the shader's mul(world, instLocal)
    on glm bytes), inverse by cofactors"""
    o2w = []
    for i, (node, _, _) in enumerate(mappings):
        w = world[node]
        o2w.append(w if inst_local is None else mat4_mul(np.asarray(inst_local[i], F), w))
    o2w = np.asarray(o2w, F)
    return o2w, mat4_inverse(o2w)
