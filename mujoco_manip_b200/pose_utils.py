"""Host-side (numpy) pose utilities with the reference's public names and conventions
(mujoco_manip/pose_utils.py:15-209): quaternions are (qx, qy, qz, qw); the 6D rotation is the first
two rows of R; 8-DOF = [x y z qx qy qz qw g], 10-DOF = [x y z r11 r12 r13 r21 r22 r23 g].

The CUDA kernels carry their own device versions of the encoders (csrc/mm_env.h); these functions
serve the single-env API (`decode_action`, action construction for expert rollouts) and the tests,
which check them against golden vectors produced by the reference's own module.
"""
from __future__ import annotations

import numpy as np


def pos_rotmat_to_se3(pos, rotmat) -> np.ndarray:
    T = np.zeros((4, 4), dtype=np.float64)
    T[3, 3] = 1.0
    T[:3, :3] = rotmat
    T[:3, 3] = pos
    return T


def se3_to_pos_rotmat(T):
    T = np.asarray(T)
    return T[:3, 3].copy(), T[:3, :3].copy()


def rotmat_to_quat_xyzw(R) -> np.ndarray:
    """Four-branch conversion keyed on the trace / largest diagonal entry; no sign canonicalisation
    (pose_utils.py:48-82)."""
    R = np.asarray(R)
    d0, d1, d2 = R[0, 0], R[1, 1], R[2, 2]
    tr = d0 + d1 + d2
    if tr > 0:
        s = 2.0 * np.sqrt(tr + 1.0)
        q = ((R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s)
    elif d0 > d1 and d0 > d2:
        s = 2.0 * np.sqrt(1.0 + d0 - d1 - d2)
        q = (0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s)
    elif d1 > d2:
        s = 2.0 * np.sqrt(1.0 + d1 - d0 - d2)
        q = ((R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s)
    else:
        s = 2.0 * np.sqrt(1.0 + d2 - d0 - d1)
        q = ((R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s)
    return np.array(q)


def quat_xyzw_to_rotmat(q) -> np.ndarray:
    """No normalisation of q; arithmetic runs in the dtype of q."""
    x, y, z, w = q
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, yz, xw, yw, zw = x * y, x * z, y * z, x * w, y * w, z * w
    return np.array([[1 - 2 * (yy + zz), 2 * (xy - zw), 2 * (xz + yw)],
                     [2 * (xy + zw), 1 - 2 * (xx + zz), 2 * (yz - xw)],
                     [2 * (xz - yw), 2 * (yz + xw), 1 - 2 * (xx + yy)]])


def rotmat_to_6d(R) -> np.ndarray:
    return np.asarray(R)[:2, :].reshape(6).astype(np.float32)


def _normalise(v) -> np.ndarray:
    return v / max(np.linalg.norm(v), 1e-12)


def rotmat_from_6d(d6) -> np.ndarray:
    d6 = np.asarray(d6)
    r1 = _normalise(d6[0:3])
    r2 = _normalise(d6[3:6] - np.dot(r1, d6[3:6]) * r1)
    return np.stack([r1, r2, np.cross(r1, r2)], axis=0)


def se3_to_pos_quat_g(T, gripper: float) -> np.ndarray:
    p, R = se3_to_pos_rotmat(T)
    return np.concatenate([p, rotmat_to_quat_xyzw(R), [gripper]]).astype(np.float32)


def se3_to_pos_rot6d_g(T, gripper: float) -> np.ndarray:
    p, R = se3_to_pos_rotmat(T)
    return np.concatenate([p, rotmat_to_6d(R), [gripper]]).astype(np.float32)


def se3_from_pos_quat_g(dof8) -> np.ndarray:
    dof8 = np.asarray(dof8)
    return pos_rotmat_to_se3(dof8[:3], quat_xyzw_to_rotmat(dof8[3:7]))


def se3_from_pos_rot6d_g(dof10) -> np.ndarray:
    dof10 = np.asarray(dof10)
    return pos_rotmat_to_se3(dof10[:3], rotmat_from_6d(dof10[3:9]))
