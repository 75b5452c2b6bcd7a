"""Host pose utilities against golden vectors produced by the reference's own pose_utils module
(tools/make_golden.py), plus the identities the reference's tests/test_pose_utils.py pins
(SE(3) round trips exact, quat round trip 1e-10, rot6d 1e-6, 8/10-DOF 1e-5, relative reconstruction 1e-4)."""
import os

import numpy as np
import pytest

from hostlib import GOLDEN
from mujoco_manip_b200 import pose_utils as P

G = np.load(os.path.join(GOLDEN, "pose_utils.npz"))


@pytest.mark.parametrize("i", range(64))
def test_against_reference_vectors(i):
    R = G["R"][i]
    np.testing.assert_array_equal(P.rotmat_to_quat_xyzw(R), G["quat"][i])  # all four branches are in the set
    np.testing.assert_array_equal(P.rotmat_to_6d(R), G["rot6d"][i])
    np.testing.assert_array_equal(P.se3_from_pos_quat_g(G["dof8"][i]), G["T_from8"][i])
    a = G["in10_T_from10"][i]
    np.testing.assert_allclose(P.se3_from_pos_rot6d_g(a[:10].astype(np.float32)).ravel(), a[10:], rtol=0, atol=1e-7)
    T = P.pos_rotmat_to_se3(G["dof8"][i][:3], R)
    np.testing.assert_array_equal(P.se3_to_pos_quat_g(T, float(G["dof8"][i][7]))[3:7], G["quat"][i].astype(np.float32))


def test_se3_round_trip_exact():
    rng = np.random.default_rng(0)
    p = rng.normal(size=3)
    R = P.quat_xyzw_to_rotmat(np.array([0.1, -0.2, 0.3, 0.9]) / np.linalg.norm([0.1, -0.2, 0.3, 0.9]))
    p2, R2 = P.se3_to_pos_rotmat(P.pos_rotmat_to_se3(p, R))
    assert np.array_equal(p, p2) and np.array_equal(R, R2)
    T = P.pos_rotmat_to_se3(p, R)
    assert T.shape == (4, 4) and np.array_equal(T[3], [0, 0, 0, 1])


def test_quat_and_rot6d_round_trips():
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        R = P.quat_xyzw_to_rotmat(q)
        q2 = P.rotmat_to_quat_xyzw(R)
        assert abs(np.linalg.norm(q2) - 1) < 1e-12
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-10
        R6 = P.rotmat_from_6d(P.rotmat_to_6d(R))
        np.testing.assert_allclose(R6, R, atol=1e-6)
        assert P.rotmat_to_6d(R).dtype == np.float32
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)


def test_dof_vectors_and_relative_reconstruction():
    rng = np.random.default_rng(2)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    T_init = P.pos_rotmat_to_se3(rng.normal(size=3), P.quat_xyzw_to_rotmat(q))
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        T = P.pos_rotmat_to_se3(rng.uniform(-1, 1, size=3), P.quat_xyzw_to_rotmat(q))
        v8, v10 = P.se3_to_pos_quat_g(T, 0.7), P.se3_to_pos_rot6d_g(T, 0.7)
        assert v8.shape == (8,) and v10.shape == (10,) and v8.dtype == np.float32 and v10.dtype == np.float32
        np.testing.assert_allclose(P.se3_from_pos_quat_g(v8), T, atol=1e-5)
        np.testing.assert_allclose(P.se3_from_pos_rot6d_g(v10), T, atol=1e-5)
        T_rel = np.linalg.inv(T_init) @ T
        np.testing.assert_allclose(T_init @ P.se3_from_pos_quat_g(P.se3_to_pos_quat_g(T_rel, 1.0)), T, atol=1e-4)
        np.testing.assert_allclose(T_init @ P.se3_from_pos_rot6d_g(P.se3_to_pos_rot6d_g(T_rel, 1.0)), T, atol=1e-4)


def test_degenerate_rot6d_is_finite():
    R = P.rotmat_from_6d(np.zeros(6))
    assert np.all(np.isfinite(R))
