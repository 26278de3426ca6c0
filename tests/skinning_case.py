"""Synthetic GPU-skinning inputs shared by the CPU and GPU tests: two skeletons over one mesh buffer laid out like
SkeletonManager does it (source attributes + overridden ranges, rend3/src/managers/skeleton.rs)."""
import numpy as np

from rend3_b200 import glam
from rend3_b200.layouts import ATTR_ABSENT, SKINNING_INPUT_DTYPE

f32 = np.float32


def build(seed=0, vertex_counts=(300, 1000), joints_per_skeleton=(4, 7)):
    rng = np.random.default_rng(seed)
    words, inputs, joint_mats, expect = [], [], [], []
    cursor = 0

    def push(arr):
        nonlocal cursor
        raw = np.ascontiguousarray(arr).view(np.uint32).reshape(-1)
        off = cursor * 4
        words.append(raw)
        cursor += len(raw)
        return off

    joint_base = 0
    for nv, nj in zip(vertex_counts, joints_per_skeleton):
        pos = rng.uniform(-1, 1, (nv, 3)).astype(f32)
        nrm = rng.standard_normal((nv, 3)).astype(f32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(f32)
        idx = rng.integers(0, nj, (nv, 4)).astype(np.uint16)
        w = rng.random((nv, 4)).astype(f32)
        w[rng.random((nv, 4)) < 0.3] = 0.0          # zero weights are skipped (skinning.wgsl:68)
        w[:, 0] = np.maximum(w[:, 0], 0.05)
        w = (w / w.sum(axis=1, keepdims=True)).astype(f32)
        rec = np.zeros((), dtype=SKINNING_INPUT_DTYPE)
        rec["base_position_offset"] = push(pos)
        rec["base_normal_offset"] = push(nrm)
        rec["base_tangent_offset"] = ATTR_ABSENT
        rec["joint_indices_offset"] = push(idx)
        rec["joint_weight_offset"] = push(w)
        rec["updated_position_offset"] = push(np.zeros((nv, 3), dtype=f32))
        rec["updated_normal_offset"] = push(np.zeros((nv, 3), dtype=f32))
        rec["updated_tangent_offset"] = ATTR_ABSENT
        rec["joint_matrix_base_offset"] = joint_base
        rec["vertex_count"] = nv
        inputs.append(rec)
        mats = []
        for _ in range(nj):
            q = rng.standard_normal(4)
            q /= np.linalg.norm(q)
            mats.append(glam.from_scale_rotation_translation(rng.uniform(0.5, 2.0, 3), q, rng.uniform(-2, 2, 3)))
        joint_mats += mats
        # float64 expectation of the skinned positions
        m64 = np.array([m.astype(np.float64).T for m in mats])   # row-major math matrices
        p4 = np.concatenate([pos.astype(np.float64), np.ones((nv, 1))], axis=1)
        acc = np.zeros((nv, 3))
        for k in range(4):
            t = np.einsum("vij,vj->vi", m64[idx[:, k]], p4)[:, :3]
            acc += t * w[:, k:k + 1].astype(np.float64)
        expect.append(acc)
        joint_base += nj
    return np.concatenate(words), np.array(inputs, dtype=SKINNING_INPUT_DTYPE), np.array(joint_mats, dtype=f32).reshape(-1, 16), expect
