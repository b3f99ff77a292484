"""Host side of the animation feed (include/b200pt.h: b200pt_set_animation / b200pt_animate).

Mirrors what SceneAnimationVk prepares for the morph / skinning compute shaders
(src/gltf_scene_animation_vk.cpp:120-260 static task data, :396-494 per-frame data): the task records and the per-frame
joint / normal matrices.  The vertex work itself runs on the device (csrc/animate.cuh); nothing here touches vertices.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class MorphTask:
    """one morphed render primitive (MorphResult + MorphGpuData, gltf_scene_animation_vk.cpp:262-330): base arrays and
    per-target deltas [numTargets, vertexCount, 3]"""
    render_prim: int
    base_positions: np.ndarray
    position_deltas: np.ndarray
    base_normals: Optional[np.ndarray] = None
    base_tangents: Optional[np.ndarray] = None
    normal_deltas: Optional[np.ndarray] = None
    tangent_deltas: Optional[np.ndarray] = None


@dataclass
class SkinTask:
    """one skinned render primitive (SkinTask + SkinGpuData, gltf_scene_animation_vk.cpp:170-226): base arrays, WEIGHTS_0 [V,4],
    JOINTS_0 [V,4] (int32), the skin's joint count"""
    render_prim: int
    base_positions: np.ndarray
    weights: np.ndarray
    joints: np.ndarray
    num_joints: int
    base_normals: Optional[np.ndarray] = None
    base_tangents: Optional[np.ndarray] = None


def joint_matrices(node_world, joint_nodes, inverse_bind, ref_node):
    """Per-frame skin matrices as cmdUpdateAnimation computes them (gltf_scene_animation_vk.cpp:470-484):
        jointMat[i]  = inverse(world[refNode]) * world[joint_i] * inverseBind[i]   (identity when the skin has fewer)
        normalMat[i] = transpose(inverse(mat3(jointMat[i])))
    node_world / inverse_bind: mathematical 4x4 matrices (row, column).  Returns (jointMats [J,4,4], normalMats [J,3,3]) in the
    glm byte order the C-ABI takes: element [j, c, r] = column c, row r.  (The reference evaluates these few products with glm in
    fp32; here they are evaluated in fp64 and rounded once -- the kernels take whatever matrices the host hands them.)"""
    inv_ref = np.linalg.inv(np.asarray(node_world[ref_node], np.float64))
    jm, nm = [], []
    for i, jn in enumerate(joint_nodes):
        ib = np.asarray(inverse_bind[i], np.float64) if i < len(inverse_bind) else np.eye(4)
        m = inv_ref @ np.asarray(node_world[jn], np.float64) @ ib
        jm.append(m.T)                               # column-major bytes
        nm.append(np.linalg.inv(m[:3, :3]))         # transpose(inverse(m3)) stored column-major: element [c, r] = inverse[c, r]
    return np.asarray(jm, np.float32), np.asarray(nm, np.float32)


def topo_levels(parents):
    """BFS levels of a node forest: (topoNodeOrder int32[numNodes], levelOffsets uint32[numLevels + 1]) -- what the reference
    uploads for world_matrix_propagate.comp.slang (one dispatch per level, parents always in an earlier level)."""
    parents = np.asarray(parents, np.int64)
    depth = np.full(len(parents), -1, np.int64)
    for n in range(len(parents)):
        chain = []
        k = n
        while depth[k] < 0 and parents[k] >= 0:
            chain.append(k)
            k = parents[k]
            if len(chain) > len(parents):
                raise ValueError("node hierarchy has a cycle")
        if depth[k] < 0:
            depth[k] = 0
        d = depth[k]
        for c in reversed(chain):
            d += 1
            depth[c] = d
    order = np.argsort(depth, kind="stable").astype(np.int32)
    counts = np.bincount(depth, minlength=int(depth.max()) + 1)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    return order, offsets
