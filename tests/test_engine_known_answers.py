"""Known-answer tests of the ENGINE layer (mj_step / mj_forward restatement) whose expected values do NOT come from the
restatement: closed forms of the published soft-constraint model (MuJoCo "Computation" chapter, restated in SURVEY.md
Appendix A4 with a worked example), conservation laws, and an independent numpy evaluation of the kinetic energy.

Both engines are held to them: the CPU oracle (oracle/) and the CUDA kernel source (host build of the very same
headers, tests/mm_emul.cpp).  The reference itself reaches this layer through libmujoco (mujoco_manip/env.py:117-121);
MuJoCo is not installable here, so these answers are what pins the layer until tools/make_golden.py --engine mujoco
is run somewhere that has it."""
import os
import re

import numpy as np
import pytest

import hostlib
from hostlib import EmulEnv, emul_forward

M_CUBE, G, H = 0.05, 9.81, 0.002
OP_FORWARD, OP_STEP = 2, 6


# ---- the published formulas, written down here independently of both engines ----
def impedance(dist, solimp=(0.9, 0.95, 0.001, 0.5, 2.0)):
    dmin, dmax, width, mid, power = solimp
    x = abs(dist) / width
    if x >= 1:
        return dmax
    y = x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    return dmin + y * (dmax - dmin)


def stiffness_damping(solref=(0.02, 1.0), dmax=0.95, h=H):
    tc, dr = max(solref[0], 2 * h), solref[1]
    return 1.0 / (dmax ** 2 * tc ** 2 * dr ** 2), 2.0 / (dmax * tc)


def pyramid_R(dist, mu, tran):
    imp = impedance(dist)
    r_first = (1 - imp) / imp * tran * (1 + mu * mu)
    return 2 * mu * mu * r_first, r_first, imp


def test_worked_example_of_the_survey():
    """cube corner on the table, penetration 1e-4 -> imp 0.901, K 2770.08, B 105.263, R_first 10.99, R_py 87.9."""
    K, B = stiffness_damping()
    R_py, R_first, imp = pyramid_R(-1e-4, 2.0, 20.0)
    assert abs(imp - 0.901) < 1e-12 and abs(K - 2770.08) < 5e-3 and abs(B - 105.263) < 5e-4
    assert abs(R_first - 10.99) < 5e-3 and abs(R_py - 87.9) < 5e-2


@pytest.fixture(scope="module")
def oracle_mod():
    from oracle import oracle

    oracle.build()
    return oracle


def _rest_state(oracle_mod):
    o = oracle_mod.OracleEnv()
    o.reset(None, 0, 0)
    for _ in range(400):
        o.mj_step()
    o.mj_forward()
    return o


@pytest.mark.parametrize("engine", ["oracle", "kernel"])
def test_cube_at_rest_force_balance_and_closed_form_penetration(oracle_mod, engine):
    """A cube resting on the table: the constraint force on it is m g, shared equally by 4 contacts x 6 pyramid rows;
    each row obeys force = D * aref with aref = K imp(d) |d|, D = 1 / R_py, R_py = 2 mu^2 (1 - imp) / imp * tran (1 + mu^2)
    (mu = 2, tran = 1 / m = 20): the engine's penetration must satisfy that closed form."""
    o = _rest_state(oracle_mod)
    K, _ = stiffness_damping()
    if engine == "oracle":
        rows = o.efc_rows()
        con = rows[rows[:, 0] == 2]
        fz = o.qfrc_constraint[[11, 17, 23]]
        dist, D, aref, force = con[:, 1], con[:, 2], con[:, 4], con[:, 5]
        assert con.shape[0] == 72
    else:
        f = emul_forward(o.qpos, o.qvel, o.ctrl, o.qacc_warmstart)
        assert f["ncon"] == 12 and np.all((f["cmeta"] >> 15) & 1)  # cube contacts are condim 4: six pyramid rows
        fz = f["fc"][[11, 17, 23]]
        dist, D, aref = np.repeat(f["cdist"], 6), np.repeat(f["cD"], 6), f["aref"].ravel()
        force = D * aref  # rows at rest: J qacc = 0
    np.testing.assert_allclose(fz, M_CUBE * G, rtol=1e-7)
    np.testing.assert_allclose(force, M_CUBE * G / 24, rtol=1e-6)
    for d, dd, a in zip(dist, D, aref):
        R_py, _, imp = pyramid_R(d, 2.0, 20.0)
        assert abs(dd - 1.0 / R_py) < 1e-9 * dd
        assert abs(a - K * imp * abs(d)) < 1e-6 * abs(a)  # velocity term B * (J v) vanishes at rest
        assert abs(dd * K * imp * abs(d) - M_CUBE * G / 24) < 1e-6 * M_CUBE * G  # closed-form equilibrium
    assert 4e-4 < -dist.max() < 6e-4  # ~0.51 mm: inside the 1 mm impedance width


def _free_fall_env():
    env = EmulEnv(1, mode="abs_pos")
    env.reset()
    q = env.st["qpos"][0]
    q[9:12] = [0.0, -0.6, 1.5]  # red cube high above the floor, away from the table
    return env


@pytest.mark.parametrize("engine", ["oracle", "kernel"])
def test_free_fall_and_free_rotation_closed_form(oracle_mod, engine):
    """No contact: semi-implicit Euler gives v_n = v_0 - g n h, z_n = z_0 + n h v_0 - g h^2 n (n + 1) / 2 exactly, linear
    momentum in x, y is conserved, and a cube spinning with constant body-frame angular velocity integrates to
    q_n = q_0 * exp(n h w / 2) (free joints integrate the quaternion with the exponential map)."""
    n = 200
    w = np.array([0.7, -1.1, 0.4])
    v0 = np.array([0.3, -0.2, 1.0])
    if engine == "oracle":
        o = oracle_mod.OracleEnv()
        o.reset(None, 0, 0)
        o.qpos[9:12] = [0.0, -0.6, 1.5]
        o.qvel[9:12] = v0
        o.qvel[12:15] = w
        for _ in range(n):
            o.mj_step()
        qpos, qvel = o.qpos.copy(), o.qvel.copy()
    else:
        env = _free_fall_env()
        env.st["qvel"][0, 9:12] = v0
        env.st["qvel"][0, 12:15] = w
        env.st["kin"][0, :9] = env.st["qpos"][0, :9]
        for _ in range(n):
            env.ops(OP_STEP)
        qpos, qvel = env.st["qpos"][0].copy(), env.st["qvel"][0].copy()
    t = n * H
    np.testing.assert_allclose(qvel[9:12], v0 + np.array([0, 0, -G * t]), atol=1e-12)
    np.testing.assert_allclose(qpos[9:11], np.array([0.0, -0.6]) + v0[:2] * t, atol=1e-12)
    assert abs(qpos[11] - (1.5 + v0[2] * t - G * H * H * n * (n + 1) / 2)) < 1e-12
    np.testing.assert_allclose(qvel[12:15], w, atol=1e-13)  # torque-free isotropic body
    ang = np.linalg.norm(w) * t
    q_exp = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / np.linalg.norm(w)])
    np.testing.assert_allclose(qpos[12:16], q_exp, atol=1e-12)


def _model_tables():
    src = open(os.path.join(hostlib.REPO, "mujoco_manip_b200", "csrc", "model_gen.h")).read()

    def arr(name, shape):
        m = re.search(r"\b" + name + r"(\[[0-9]+\])+\s*=\s*\{(.*?)\};", src, re.S)
        vals = [float(x) for x in re.findall(r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?", m.group(2))]
        return np.array(vals).reshape(shape)

    return arr("mm_body_mass", (19,)), arr("mm_body_inertia", (19, 3, 3))


def test_mass_matrix_matches_kinetic_energy_from_finite_differences(oracle_mod):
    """v^T M v / 2 (CRBA, both engines) against the kinetic energy computed from its definition: body COM velocities and
    angular velocities obtained by central differences of the forward kinematics, with the masses and inertia tensors of
    the model tables - an evaluation that shares no code with the mass-matrix algorithms."""
    mass, inertia = _model_tables()
    o = oracle_mod.OracleEnv()
    o.reset(None, 0, 0)
    rng = np.random.default_rng(3)
    q0 = o.qpos.copy()
    q0[:7] += rng.uniform(-0.3, 0.3, size=7)
    eps = 1e-6
    for trial in range(4):
        v = np.zeros(27)
        v[:9] = rng.uniform(-1, 1, size=9) * np.array([1, 1, 1, 1, 1, 1, 1, 0.02, 0.02])

        def kin(sign):
            o.qpos[:] = q0
            o.qpos[:9] += sign * eps * v[:9]
            o.mj_forward()
            return o.xipos.copy(), o.xmat.copy().reshape(19, 3, 3)

        (pp, Rp), (pm, Rm) = kin(+1), kin(-1)
        o.qpos[:] = q0
        o.mj_forward()
        R0 = o.xmat.copy().reshape(19, 3, 3)
        T = 0.0
        for b in range(2, 12):  # link1 .. right_finger (link0 is static)
            vc = (pp[b] - pm[b]) / (2 * eps)
            dR = (Rp[b] - Rm[b]) / (2 * eps) @ R0[b].T  # [w]x
            wv = np.array([dR[2, 1], dR[0, 2], dR[1, 0]])
            Iw = R0[b] @ inertia[b] @ R0[b].T  # the tables hold the full inertia tensor about the COM in the body frame
            T += 0.5 * mass[b] * vc @ vc + 0.5 * wv @ Iw @ wv
        arm = np.array([0.1] * 9)  # joint armature adds 0.1 v_i^2 / 2 per robot dof (panda.xml:9)
        T += 0.5 * np.sum(arm * v[:9] ** 2)
        M = o.M.reshape(27, 27)
        T_crba = 0.5 * v @ M @ v
        f = emul_forward(q0, np.zeros(27), o.ctrl)
        T_kernel = 0.5 * v[:9] @ f["Mr"] @ v[:9]
        assert abs(T_crba - T) < 2e-6 * max(T, 1e-3), (trial, T_crba, T)
        assert abs(T_kernel - T_crba) < 1e-12 * max(T, 1e-3)


@pytest.mark.parametrize("engine", ["oracle", "kernel"])
def test_gripper_servo_and_equality_steady_state(oracle_mod, engine):
    """Gripper actuator on the tendon 0.5 (q7 + q8) with gain 0.01568627 and bias (0, -100, -10) (panda.xml:276-277): at
    rest the tendon length is gain * ctrl / 100; the joint equality f1 = f2 (panda.xml:260-262) keeps the fingers equal.
    ctrl = 127.5 -> both fingers at 0.02 m."""
    if engine == "oracle":
        o = oracle_mod.OracleEnv()
        o.reset(None, 0, 0)
        o.ctrl[7] = 127.5
        for _ in range(1500):
            o.mj_step()
        q = o.qpos.copy()
        v = o.qvel.copy()
    else:
        env = EmulEnv(1, mode="abs_pos")
        env.reset()
        env.st["ctrl"][0, 7] = 127.5
        for _ in range(1500):
            env.ops(OP_STEP)
        q, v = env.st["qpos"][0].copy(), env.st["qvel"][0].copy()
    assert abs(q[7] - q[8]) < 1e-7
    assert abs(0.5 * (q[7] + q[8]) - 0.01568627451 * 127.5 / 100) < 1e-6
    assert np.abs(v[7:9]).max() < 1e-8
