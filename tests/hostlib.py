"""Test helpers: state allocation + ctypes access to the 1-lane host emulation of the kernel source
(tests/_build/libmm_emul.so, built from tests/mm_emul.cpp + the kernel headers with g++).  Test-only."""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
_EMUL = os.path.join(REPO, "tests", "_build", "libmm_emul.so")
MODES = ("abs_pos", "ee_pos_quat_g", "ee_pos_rot6d_g", "ee_pos_quat_g_rel", "ee_pos_rot6d_g_rel")
ACTION_DIMS = {"abs_pos": 4, "ee_pos_quat_g": 8, "ee_pos_rot6d_g": 10, "ee_pos_quat_g_rel": 8, "ee_pos_rot6d_g_rel": 10}
REWARDS = ("dense", "sparse", "staged")

STATE_FIELDS = [("qpos", 30, np.float64), ("qvel", 27, np.float64), ("ctrl", 8, np.float64), ("warm", 27, np.float64),
                ("tinit", 12, np.float64), ("eepose", 12, np.float64), ("fsm_f", 6, np.float64), ("hwm", 5, np.float64),
                ("kin", 18, np.float64),
                ("step_count", 1, np.int32), ("task", 2, np.int32), ("fsm_i", 5, np.int32), ("fsm_tasks", 20, np.int32), ("flags", 1, np.int32),
                ("diag", 4, np.int32)]


def build_emul(tag="", defines=()):
    """Host build of the kernel source; `tag` / `defines` give a second build with extra -D switches (test hooks)."""
    out = _EMUL if not tag else _EMUL.replace(".so", f"_{tag}.so")
    srcdir = os.path.join(REPO, "mujoco_manip_b200", "csrc")
    srcs = [os.path.join(srcdir, f) for f in os.listdir(srcdir) if f.endswith(".h")] + [os.path.join(REPO, "tests", "mm_emul.cpp")]
    if os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in srcs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", *defines, "-o", out,
                           os.path.join(REPO, "tests", "mm_emul.cpp")])
    return out


class EmulEnv:
    """N environments stepped by the host build (G = 1) of the kernel source."""

    def __init__(self, n, mode="ee_pos_quat_g_rel", reward="dense", max_steps=500, use_float=False, lib=None):
        self.L = C.CDLL(lib or build_emul())
        self.n, self.mode, self.reward, self.max_steps, self.use_float = n, mode, reward, max_steps, int(use_float)
        self.st = {k: np.zeros((n, d), dtype=t) for k, d, t in STATE_FIELDS}
        self._sp = (C.c_void_p * len(STATE_FIELDS))(*[self.st[k].ctypes.data for k, _, _ in STATE_FIELDS])
        self.obs = np.zeros((n, 85), dtype=np.float32)
        self.reward_buf = np.zeros(n, dtype=np.float32)
        self.term = np.zeros(n, dtype=np.uint8)
        self.trunc = np.zeros(n, dtype=np.uint8)
        self.succ = np.zeros(n, dtype=np.uint8)
        self.rc = np.zeros((n, 6), dtype=np.float32)
        self.tgt = np.zeros((n, 4), dtype=np.float32)
        self._op = (C.c_void_p * 6)(self.obs.ctypes.data, self.reward_buf.ctypes.data, self.term.ctypes.data,
                                    self.trunc.ctypes.data, self.succ.ctypes.data, self.rc.ctypes.data)

    def reset(self, obj_xy=None, task=None, mask=None, yaw=None):
        task = np.ascontiguousarray(np.zeros((self.n, 2)) if task is None else task, dtype=np.int32)
        xy = None if obj_xy is None else np.ascontiguousarray(obj_xy, dtype=np.float64).reshape(self.n, 6)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        cs = None
        if yaw is not None:  # [n,3] angles -> (cos, sin)(theta / 2)
            th = np.asarray(yaw, dtype=np.float64).reshape(self.n, 3)
            cs = np.ascontiguousarray(np.stack([np.cos(th / 2), np.sin(th / 2)], axis=2).reshape(self.n, 6))
        self.L.emul_reset_yaw(self.n, self._sp, None if m is None else m.ctypes.data_as(C.c_void_p),
                              None if xy is None else xy.ctypes.data_as(C.c_void_p),
                              None if cs is None else cs.ctypes.data_as(C.c_void_p), task.ctypes.data_as(C.c_void_p),
                              self.obs.ctypes.data_as(C.c_void_p), self.tgt.ctypes.data_as(C.c_void_p), self.use_float)
        return self.obs.copy()

    def step(self, actions):
        a = np.zeros((self.n, 10), dtype=np.float32)
        act = np.asarray(actions, dtype=np.float32).reshape(self.n, -1)
        a[:, : act.shape[1]] = act
        self.L.emul_step(self.n, self._sp, a.ctypes.data_as(C.c_void_p), MODES.index(self.mode), REWARDS.index(self.reward),
                         self.max_steps, self._op, self.tgt.ctypes.data_as(C.c_void_p), self.use_float)
        return self.obs.copy(), self.reward_buf.copy(), self.term.astype(bool), self.trunc.astype(bool), self.succ.astype(bool)

    def step_counted(self, actions):
        """Same step executed with the operation-counting scalar: returns (flops[3] = stage A, convex stage, stage C)
        summed over the envs; the state advances exactly as in `step` (FP64)."""
        a = np.zeros((self.n, 10), dtype=np.float32)
        act = np.asarray(actions, dtype=np.float32).reshape(self.n, -1)
        a[:, : act.shape[1]] = act
        fl = np.zeros(3, dtype=np.int64)
        self.L.emul_step_counted(self.n, self._sp, a.ctypes.data_as(C.c_void_p), MODES.index(self.mode), REWARDS.index(self.reward),
                                 self.max_steps, self._op, self.tgt.ctypes.data_as(C.c_void_p), fl.ctypes.data_as(C.c_void_p))
        return fl

    def ops(self, ops, target=None):
        t = None if target is None else np.ascontiguousarray(target, dtype=np.float64).reshape(self.n, 3)
        self.L.emul_ops(self.n, self._sp, int(ops), None if t is None else t.ctypes.data_as(C.c_void_p), self.use_float)

    def fsm_plan(self, n_steps=16):
        a = np.zeros((self.n, 10), dtype=np.float32)
        self.L.emul_fsm_plan(self.n, self._sp, n_steps, a.ctypes.data_as(C.c_void_p))
        return a[:, :4].copy()


def reltol(a, b, tol):
    """|a-b| <= tol * max(|b|, 1) elementwise."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))


def emul_forward(qpos, qvel, ctrl, warm=None, maxcon=256):
    """One forward pass of the kernel source (FP64, host build) on a raw state with the intermediates exposed."""
    L = C.CDLL(build_emul())
    qpos = np.ascontiguousarray(qpos, dtype=np.float64)
    qvel = np.ascontiguousarray(qvel, dtype=np.float64)
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    warm = np.zeros(27) if warm is None else np.ascontiguousarray(warm, dtype=np.float64)
    out = dict(Mr=np.zeros((9, 9)), fs=np.zeros(27), qacc_smooth=np.zeros(27), qacc=np.zeros(27), fc=np.zeros(27),
               bpos=np.zeros((13, 3)), bR=np.zeros((13, 9)), cpos=np.zeros((maxcon, 3)), cn=np.zeros((maxcon, 3)),
               cdist=np.zeros(maxcon), cmeta=np.zeros(maxcon, dtype=np.int32), cD=np.zeros(maxcon), aref=np.zeros((maxcon, 6)))
    ncon, niter = C.c_int(), C.c_int()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.emul_forward_debug(p(qpos), p(qvel), p(ctrl), p(warm), p(out["Mr"]), p(out["fs"]), p(out["qacc_smooth"]), p(out["qacc"]),
                         p(out["fc"]), p(out["bpos"]), p(out["bR"]), C.byref(ncon), p(out["cpos"]), p(out["cn"]), p(out["cdist"]),
                         p(out["cmeta"]), C.byref(niter), p(out["cD"]), p(out["aref"]))
    n = ncon.value
    for k in ("cpos", "cn", "cdist", "cmeta", "cD", "aref"):
        out[k] = out[k][:n]
    out["ncon"], out["niter"] = n, niter.value
    return out
