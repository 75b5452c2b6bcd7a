"""ORACLE - TEST INFRASTRUCTURE ONLY.

Stand-ins for the two third-party modules the reference imports but this image does not have:

* ``mujoco``    - exactly the surface the reference touches (SURVEY 8b "native boundary today"),
                  backed by the CPU oracle in ``oracle/engine.cpp``;
* ``gymnasium`` - ``Env`` / ``spaces.Box`` / ``spaces.Dict`` / ``register``, with
                  gymnasium's seeding rule ``np_random = default_rng(seed)``.

``install()`` puts them in ``sys.modules`` so that the UNMODIFIED reference package under
``/root/reference`` imports and runs in this container.  That is how the oracle's Python-layer
restatement (oracle/hotpath.cpp) and the CUDA path are pinned against the reference's own code:
``tools/make_golden.py`` runs the reference on top of this shim and commits the outputs under
``tests/golden/``.  Rendering is stubbed (black images) - images are out of scope.
"""
from __future__ import annotations

import enum
import os
import sys
import tempfile
import types

import numpy as np

from . import oracle as _o


class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0
    mjOBJ_BODY = 1
    mjOBJ_JOINT = 3
    mjOBJ_GEOM = 5
    mjOBJ_CAMERA = 7
    mjOBJ_KEY = 21


_CAMERAS = ["overhead", "side", "wrist"]
_KEYS = ["home", "scene_start"]


class _Opt:
    timestep = 0.002


class MjModel:
    def __init__(self, n_cam=2):
        L = _o.lib()
        self.nq = L.orc_model_int(b"nq")
        self.nv = L.orc_model_int(b"nv")
        self.nu = L.orc_model_int(b"nu")
        self.nbody = L.orc_model_int(b"nbody")
        self.ngeom = L.orc_model_int(b"ngeom")
        self.njnt = L.orc_model_int(b"njnt")
        self.ncam = n_cam
        self.opt = _Opt()
        self.body_names = [L.orc_body_name(i).decode() for i in range(self.nbody)]
        self.jnt_names = [L.orc_jnt_name(i).decode() for i in range(self.njnt)]
        self.geom_bodyid = np.array([L.orc_geom_bodyid(i) for i in range(self.ngeom)], dtype=np.int32)
        self.jnt_qposadr = np.array([L.orc_jnt_qposadr(i) for i in range(self.njnt)], dtype=np.int32)
        rng = np.zeros((self.njnt, 2))
        for i in range(self.njnt):
            r = np.zeros(2)
            L.orc_jnt_range(i, r.ctypes.data)
            rng[i] = r
        self.jnt_range = rng
        self.cam_fovy = np.array([45.0, 45.0, 128.0][:n_cam])

    @staticmethod
    def from_xml_path(path):
        _check_scene(path)
        return MjModel(n_cam=2)


def _check_scene(path):
    with open(path) as f:
        txt = f.read()
    if 'model="pick_and_place_scene"' not in txt:
        raise ValueError("fake mujoco only knows the pick_and_place_scene model")


class _Contact:
    def __init__(self, g1, g2, dist):
        self.geom1, self.geom2, self.dist = g1, g2, dist


class _ContactList:
    def __init__(self, env):
        self._env = env

    def __getitem__(self, i):
        c = self._env.contacts()[i]
        return _Contact(c["geom1"], c["geom2"], c["dist"])


class MjData:
    def __init__(self, model):
        self._model = model
        self._env = _o.OracleEnv()
        e = self._env
        self.qpos, self.qvel, self.ctrl = e.qpos, e.qvel, e.ctrl
        self.xpos, self.xmat = e.xpos, e.xmat
        self.cam_xpos, self.cam_xmat = e.cam_xpos, e.cam_xmat
        self.qacc, self.qacc_warmstart = e.qacc, e.qacc_warmstart
        self.contact = _ContactList(e)
        # mj_makeData leaves the state at qpos0 (not the keyframe)
        e.qpos[:9] = 0
        e.qvel[:] = 0
        e.ctrl[:] = 0
        e.qacc_warmstart[:] = 0

    @property
    def ncon(self):
        return self._env.ncon

    @property
    def time(self):
        return float(self._env.time[0])


def mj_name2id(model, objtype, name):
    try:
        if objtype == mjtObj.mjOBJ_BODY:
            return model.body_names.index(name)
        if objtype == mjtObj.mjOBJ_JOINT:
            return model.jnt_names.index(name)
        if objtype == mjtObj.mjOBJ_CAMERA:
            return _CAMERAS[: model.ncam].index(name)
        if objtype == mjtObj.mjOBJ_KEY:
            return _KEYS.index(name)
    except ValueError:
        return -1
    return -1


def mj_id2name(model, objtype, idx):
    if objtype == mjtObj.mjOBJ_BODY:
        return model.body_names[idx]
    if objtype == mjtObj.mjOBJ_JOINT:
        return model.jnt_names[idx]
    return None


def mj_resetDataKeyframe(model, data, key):
    if key != 1:
        raise NotImplementedError("only the scene_start keyframe is modelled")
    data._env.mj_reset_keyframe()


def mj_forward(model, data):
    data._env.mj_forward()


def mj_step(model, data):
    data._env.mj_step()


def mj_jac(model, data, jacp, jacr, point, body):
    jp, jr = data._env.mj_jac(point, body)
    if jacp is not None:
        jacp[:] = jp
    if jacr is not None:
        jacr[:] = jr


class _SpecCamera:
    name = ""
    pos = None
    quat = None
    fovy = 45.0


class _SpecBody:
    def __init__(self, spec):
        self._spec = spec

    def add_camera(self):
        c = _SpecCamera()
        self._spec._cams.append(c)
        return c


class MjSpec:
    def __init__(self):
        self._cams = []

    @staticmethod
    def from_file(path):
        _check_scene(path)
        return MjSpec()

    def body(self, name):
        assert name == "hand"
        return _SpecBody(self)

    def compile(self):
        for c in self._cams:  # the oracle hard-codes the wrist camera of env.py:57-64
            assert c.name == "wrist" and abs(c.fovy - 128.0) < 1e-12
            assert np.allclose(c.pos, [-0.07, 0.0, 0.055]) and np.allclose(c.quat, [-0.0616, -0.7044, 0.7044, 0.0616])
        return MjModel(n_cam=2 + len(self._cams))


class Renderer:
    def __init__(self, model, height=224, width=224):
        self._h, self._w = height, width

    def update_scene(self, data, camera=None):
        pass

    def render(self):
        return np.zeros((self._h, self._w, 3), dtype=np.uint8)

    def close(self):
        pass


# ---- gymnasium shim -----------------------------------------------------------------------------
class _Space:
    pass


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low) if np.ndim(low) else np.shape(high)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        out = np.empty(self.shape, dtype=np.float64)
        lo, hi = self.low.astype(np.float64), self.high.astype(np.float64)
        unb = np.isinf(lo) & np.isinf(hi)
        bnd = ~np.isinf(lo) & ~np.isinf(hi)
        out[unb] = self._rng.normal(size=unb.sum())
        out[bnd] = self._rng.uniform(lo[bnd], hi[bnd])
        return out.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Dict(_Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()


class Env:
    metadata: dict = {}
    render_mode = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)

    @property
    def unwrapped(self):
        return self


_registry = {}


def register(id, entry_point=None, **kwargs):
    _registry[id] = entry_point


_REDIRECT = os.path.join(tempfile.gettempdir(), "fake_mujoco_scene_tmp")


def install(reference_root="/root/reference"):
    """Install the shims and make the reference importable (read-only: temp files are redirected)."""
    if "mujoco" in sys.modules and getattr(sys.modules["mujoco"], "__fake__", False):
        return
    m = types.ModuleType("mujoco")
    m.__fake__ = True
    for k in ("mjtObj", "MjModel", "MjData", "MjSpec", "Renderer", "mj_name2id", "mj_id2name",
              "mj_resetDataKeyframe", "mj_forward", "mj_step", "mj_jac"):
        setattr(m, k, globals()[k])
    v = types.ModuleType("mujoco.viewer")
    v.launch_passive = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no viewer in the fake"))
    m.viewer = v
    sys.modules["mujoco"] = m
    sys.modules["mujoco.viewer"] = v

    g = types.ModuleType("gymnasium")
    s = types.ModuleType("gymnasium.spaces")
    s.Box, s.Dict, s.Space = Box, Dict, _Space
    g.spaces, g.Env, g.register = s, Env, register
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = register
    envs.registration = reg
    g.envs = envs
    sys.modules.update({"gymnasium": g, "gymnasium.spaces": s, "gymnasium.envs": envs,
                        "gymnasium.envs.registration": reg})

    # env.py:44 writes a temp copy of the scene next to panda.xml; /root/reference is read-only
    os.makedirs(_REDIRECT, exist_ok=True)
    real_mkstemp = tempfile.mkstemp

    def mkstemp(suffix=None, prefix=None, dir=None, text=False):
        if dir is not None and os.path.abspath(dir).startswith(os.path.abspath(reference_root)):
            dir = _REDIRECT
        return real_mkstemp(suffix=suffix, prefix=prefix, dir=dir, text=text)

    tempfile.mkstemp = mkstemp
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
