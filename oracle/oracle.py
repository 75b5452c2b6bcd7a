"""ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (mujoco_manip_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ACTION_MODES = ("abs_pos", "ee_pos_quat_g", "ee_pos_rot6d_g", "ee_pos_quat_g_rel", "ee_pos_rot6d_g_rel")
REWARD_TYPES = ("dense", "sparse", "staged")
ACTION_DIMS = {"abs_pos": 4, "ee_pos_quat_g": 8, "ee_pos_rot6d_g": 10, "ee_pos_quat_g_rel": 8, "ee_pos_rot6d_g_rel": 10}
OBS_FULL_DIM = 85


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no GPU needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("engine.cpp", "hotpath.cpp", "api.cpp", "engine.h", "hotpath.h", "ccd.h")]
    srcs.append(os.path.join(_HERE, "..", "mujoco_manip_b200", "csrc", "model_gen.h"))
    if not force and os.path.exists(_LIB_PATH):
        t = os.path.getmtime(_LIB_PATH)
        if all(os.path.getmtime(s) <= t for s in srcs):
            return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_new.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_ptr.restype = C.POINTER(C.c_double)
        L.orc_ptr.argtypes = [C.c_void_p, C.c_char_p]
        for n in ("orc_ncon", "orc_nefc", "orc_niter"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = C.c_int
        L.orc_efc.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_contact.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                  C.c_void_p, C.c_void_p]
        for n in ("orc_mj_step", "orc_mj_forward", "orc_mj_kinematics", "orc_mj_reset_keyframe", "orc_fsm_reset",
                  "orc_stats_clear"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_mj_jac.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_env_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_env_reset_yaw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_env_step.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_env_obs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_ik.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fsm_plan.argtypes = [C.c_void_p, C.c_int]
        L.orc_fsm_action.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_fsm_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_model_int.argtypes = [C.c_char_p]
        for n in ("orc_body_name", "orc_geom_name", "orc_jnt_name"):
            getattr(L, n).argtypes = [C.c_int]
            getattr(L, n).restype = C.c_char_p
        L.orc_geom_bodyid.argtypes = [C.c_int]
        L.orc_jnt_qposadr.argtypes = [C.c_int]
        L.orc_jnt_range.argtypes = [C.c_int, C.c_void_p]
        L.orc_run_fsm_episode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_run_fsm_episode.restype = C.c_int
        L.orc_bench_random.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int]
        L.orc_bench_random.restype = C.c_double
        L.orc_bench_random2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_void_p]
        L.orc_bench_random2.restype = C.c_double
        _lib = L
    return _lib


class OracleEnv:
    """One CPU environment: the engine state plus the reference's Python-layer state."""

    _SHAPES = {"qpos": (30,), "qvel": (27,), "ctrl": (8,), "qacc": (27,), "qacc_warmstart": (27,),
               "qacc_smooth": (27,), "qfrc_bias": (27,), "qfrc_smooth": (27,), "qfrc_constraint": (27,),
               "qfrc_actuator": (27,), "xpos": (19, 3), "xmat": (19, 9), "xquat": (19, 4), "xipos": (19, 3),
               "geom_xpos": (47, 3), "geom_xmat": (47, 9), "cam_xpos": (3, 3), "cam_xmat": (3, 9), "M": (27, 27),
               "time": (1,), "init_pos": (3,), "init_R": (3, 3)}

    def __init__(self, action_mode="ee_pos_quat_g_rel", reward_type="dense", max_episode_steps=500, flags=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_new())
        self.action_mode = action_mode
        self.L.orc_config(self.h, ACTION_MODES.index(action_mode), REWARD_TYPES.index(reward_type), max_episode_steps, flags)
        for k, shp in self._SHAPES.items():
            p = self.L.orc_ptr(self.h, k.encode())
            setattr(self, k, np.ctypeslib.as_array(p, shape=(int(np.prod(shp)),)).reshape(shp))

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    # engine level -------------------------------------------------------------------------
    def mj_step(self):
        self.L.orc_mj_step(self.h)

    def mj_forward(self):
        self.L.orc_mj_forward(self.h)

    def mj_reset_keyframe(self):
        self.L.orc_mj_reset_keyframe(self.h)

    def mj_jac(self, point, body):
        jp = np.zeros((3, 27))
        jr = np.zeros((3, 27))
        pt = np.ascontiguousarray(point, dtype=np.float64)
        self.L.orc_mj_jac(self.h, jp.ctypes.data, jr.ctypes.data, pt.ctypes.data, int(body))
        return jp, jr

    @property
    def ncon(self):
        return self.L.orc_ncon(self.h)

    @property
    def nefc(self):
        return self.L.orc_nefc(self.h)

    @property
    def niter(self):
        return self.L.orc_niter(self.h)

    def efc_rows(self):
        """Constraint rows of the last forward pass: array [nefc, 6] = type, pos, D, R, aref, force."""
        rows = []
        buf = np.zeros(5)
        for i in range(self.nefc):
            t = self.L.orc_efc(self.h, i, buf.ctypes.data)
            rows.append([t, *buf])
        return np.array(rows).reshape(-1, 6)

    def contacts(self):
        out = []
        for i in range(self.ncon):
            g1, g2, dist = C.c_int(), C.c_int(), C.c_double()
            pos = np.zeros(3)
            frame = np.zeros(9)
            self.L.orc_contact(self.h, i, C.byref(g1), C.byref(g2), C.byref(dist), pos.ctypes.data, frame.ctypes.data)
            out.append(dict(geom1=g1.value, geom2=g2.value, dist=dist.value, pos=pos, frame=frame.reshape(3, 3)))
        return out

    # hot path -----------------------------------------------------------------------------
    def reset(self, obj_xy=None, obj_idx=0, bin_idx=0, yaw=None):
        """yaw: three angles theta (randomize_yaw=True of randomization.py:55-62) or None."""
        p = None
        if obj_xy is not None:
            xy = np.ascontiguousarray(obj_xy, dtype=np.float64).reshape(6)
            p = xy.ctypes.data
        if yaw is not None:
            th = np.asarray(yaw, dtype=np.float64).reshape(3)
            cs = np.ascontiguousarray(np.stack([np.cos(th / 2), np.sin(th / 2)], axis=1).reshape(6))
            self.L.orc_env_reset_yaw(self.h, p, cs.ctypes.data, int(obj_idx), int(bin_idx))
        else:
            self.L.orc_env_reset(self.h, p, int(obj_idx), int(bin_idx))
        return self.obs()

    def obs(self):
        o = np.zeros(OBS_FULL_DIM, dtype=np.float32)
        self.L.orc_env_obs(self.h, o.ctypes.data)
        return o

    def step(self, action):
        a = np.zeros(10, dtype=np.float32)
        act = np.asarray(action, dtype=np.float32)
        a[: act.size] = act
        o = np.zeros(OBS_FULL_DIM, dtype=np.float32)
        r = C.c_double()
        te, tr, su = C.c_int(), C.c_int(), C.c_int()
        rc = np.zeros(6, dtype=np.float32)
        self.L.orc_env_step(self.h, a.ctypes.data, o.ctypes.data, C.addressof(r), C.addressof(te), C.addressof(tr),
                            C.addressof(su), rc.ctypes.data)
        return o, r.value, bool(te.value), bool(tr.value), dict(success=bool(su.value), reward_components=rc)

    def ik(self, target):
        t = np.ascontiguousarray(target, dtype=np.float64)
        q = np.zeros(7)
        self.L.orc_ik(self.h, t.ctypes.data, q.ctypes.data)
        return q

    def decode(self, action):
        a = np.zeros(10, dtype=np.float32)
        act = np.asarray(action, dtype=np.float32)
        a[: act.size] = act
        t = np.zeros(3)
        g = C.c_float()
        self.L.orc_decode(self.h, a.ctypes.data, t.ctypes.data, C.addressof(g))
        return t, g.value

    def fsm_reset(self):
        self.L.orc_fsm_reset(self.h)

    def fsm_plan(self, n=16):
        self.L.orc_fsm_plan(self.h, int(n))

    def fsm_action(self):
        a = np.zeros(4, dtype=np.float32)
        self.L.orc_fsm_action(self.h, a.ctypes.data)
        return a

    def fsm_get(self):
        st = np.zeros(5, dtype=np.int32)
        t = np.zeros(3)
        te = np.zeros(3)
        self.L.orc_fsm_get(self.h, st.ctypes.data, t.ctypes.data, te.ctypes.data)
        return dict(state=int(st[0]), task_index=int(st[1]), counter=int(st[2]), gripper_open=int(st[3]),
                    has_target=int(st[4]), target=t, transit_end=te)

    def stats(self):
        s = np.zeros(13, dtype=np.int64)
        self.L.orc_stats(self.h, s.ctypes.data)
        return dict(zip(STAT_KEYS, s.tolist()))

    def stats_clear(self):
        self.L.orc_stats_clear(self.h)

    def run_fsm_episode(self, obj_xy=None, obj_idx=0, bin_idx=0, max_steps=2000):
        p = None
        if obj_xy is not None:
            xy = np.ascontiguousarray(obj_xy, dtype=np.float64).reshape(6)
            p = xy.ctypes.data
        n = C.c_int()
        hist = np.zeros(12, dtype=np.int32)
        s = self.L.orc_run_fsm_episode(self.h, p, int(obj_idx), int(bin_idx), int(max_steps), C.addressof(n), hist.ctypes.data)
        return bool(s), n.value, hist


STAT_KEYS = ("substeps", "ncon", "nefc", "newton_iters", "ls_evals", "narrow_tests", "ccd_tests", "max_ncon", "max_nefc",
             "max_newton", "forwards", "nnzJ", "nnzJ2")


def bench_random_stats(n_envs, n_steps, mode="ee_pos_quat_g_rel", seed=1234, nthreads=1, flags=0):
    """-> (env-steps/s, work counters summed over the sample) - feeds bench.py's FLOP model."""
    s = np.zeros(13, dtype=np.int64)
    v = lib().orc_bench_random2(n_envs, n_steps, ACTION_MODES.index(mode), seed, nthreads, flags, s.ctypes.data)
    return v, dict(zip(STAT_KEYS, s.tolist()))


def bench_random(n_envs, n_steps, mode="ee_pos_quat_g_rel", seed=1234, nthreads=1, flags=0):
    return lib().orc_bench_random(n_envs, n_steps, ACTION_MODES.index(mode), seed, nthreads, flags)


def sample_placement(seed, x_range=(-0.20, 0.20), y_range=(0.30, 0.45), min_sep=0.08, rng=None):
    """numpy-PCG64 rejection sampler, same draw order as randomization.py:70-98."""
    rng = rng if rng is not None else np.random.default_rng(seed)
    for _ in range(1000):
        xs = rng.uniform(x_range[0], x_range[1], size=3)
        ys = rng.uniform(y_range[0], y_range[1], size=3)
        ok = True
        for i in range(3):
            for j in range(i + 1, 3):
                dx, dy = xs[i] - xs[j], ys[i] - ys[j]
                if dx * dx + dy * dy < min_sep * min_sep:
                    ok = False
        if ok:
            return np.stack([xs, ys], axis=1), rng
    raise RuntimeError("Failed to sample 3 positions")
