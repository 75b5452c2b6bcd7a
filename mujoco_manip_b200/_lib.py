"""ctypes binding of the C-ABI CUDA library (include/mm_manip.h, built from csrc/mm_kernels.cu).

The library is the ONLY compute path of this package.  There is no CPU fallback: if the shared
object is missing or no CUDA device is visible, construction fails with an error.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("MM_LIB_PATH") or os.path.join(_HERE, "_C", "libmm_manip.so")  # MM_LIB_PATH: A/B builds
SRC_DIR = os.path.join(_HERE, "csrc")

NQ, NV, NU = 30, 27, 8
OBS_DIM = 85
ACTION_STRIDE = 10
ACTION_MODES = ("abs_pos", "ee_pos_quat_g", "ee_pos_rot6d_g", "ee_pos_quat_g_rel", "ee_pos_rot6d_g_rel")
ACTION_DIMS = {"abs_pos": 4, "ee_pos_quat_g": 8, "ee_pos_rot6d_g": 10, "ee_pos_quat_g_rel": 8, "ee_pos_rot6d_g_rel": 10}
REWARD_TYPES = ("dense", "sparse", "staged")

# every symbol include/mm_manip.h declares (tests check the built library exports each one)
EXPORTS = ("mm_create", "mm_destroy", "mm_last_error", "mm_workspace_bytes", "mm_reset", "mm_step", "mm_step_host", "mm_step_host_async",
           "mm_fsm_plan", "mm_launch_count", "mm_sample_placements", "mm_measure_fma_peak", "mm_set_cycle_buffer", "mm_ops", "mm_set_schedule", "mm_expert_actions",
           "mm_set_placement_yaw", "mm_sample_yaw", "mm_sample_episode", "mm_post_step", "mm_stage_timing", "mm_stage_times", "mm_host_staging")


class MMConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32), ("group", C.c_int32),
                ("reward_type", C.c_int32), ("max_episode_steps", C.c_int32)]


# (name, width, is_double) in the order of struct mm_state
STATE_FIELDS = (("qpos", 30, True), ("qvel", 27, True), ("ctrl", 8, True), ("warm", 27, True), ("tinit", 12, True),
                ("eepose", 12, True), ("fsm_f", 6, True), ("hwm", 5, True), ("kin", 18, True), ("step_count", 1, False),
                ("task", 2, False), ("fsm_i", 5, False), ("fsm_tasks", 20, False),
                ("flags", 1, False), ("diag", 4, False))


class MMState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _, _ in STATE_FIELDS]


class MMStepOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p), ("truncated", C.c_void_p),
                ("success", C.c_void_p), ("reward_components", C.c_void_p)]


NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-diag-suppress=170,128"]  # 170: the EPA workspace deliberately spans consecutive members of the scratch struct
BUILD_DIR = os.path.join(_HERE, "_C", "obj")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into _C/libmm_manip.so (nvcc cross-compiles without a GPU).
    One translation unit per (precision, group) kernel instantiation, compiled in parallel."""
    from concurrent.futures import ThreadPoolExecutor

    hdrs = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(REPO, "include", "mm_manip.h"))
    units = sorted(f for f in os.listdir(SRC_DIR) if f.endswith(".cu"))
    os.makedirs(BUILD_DIR, exist_ok=True)
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)

    def compile_one(u):
        src, obj = os.path.join(SRC_DIR, u), os.path.join(BUILD_DIR, u[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(newest_hdr, os.path.getmtime(src)):
            return obj, False
        extra = os.environ.get("MM_NVCC_EXTRA", "").split()
        cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        subprocess.check_call(cmd)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_one, units))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(LIB_PATH):
        subprocess.check_call(["nvcc", "--shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(this package has no CPU path)")
    L = C.CDLL(LIB_PATH)
    L.mm_last_error.restype = C.c_char_p
    L.mm_workspace_bytes.restype = C.c_size_t
    L.mm_workspace_bytes.argtypes = [C.POINTER(MMConfig)]
    L.mm_create.argtypes = [C.POINTER(MMConfig), C.POINTER(C.c_void_p)]
    L.mm_destroy.argtypes = [C.c_void_p]
    L.mm_destroy.restype = None
    L.mm_reset.argtypes = [C.c_void_p, C.POINTER(MMState), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mm_step.argtypes = [C.c_void_p, C.POINTER(MMState), C.c_void_p, C.c_int, C.POINTER(MMStepOut), C.c_void_p]
    L.mm_step_host.argtypes = [C.c_void_p, C.POINTER(MMState), C.c_void_p, C.c_int] + [C.c_void_p] * 6
    L.mm_step_host_async.argtypes = L.mm_step_host.argtypes
    L.mm_fsm_plan.argtypes = [C.c_void_p, C.POINTER(MMState), C.c_int, C.c_void_p, C.c_void_p]
    L.mm_sample_placements.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mm_measure_fma_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.mm_ops.argtypes = [C.c_void_p, C.POINTER(MMState), C.c_int, C.c_void_p, C.c_void_p]
    L.mm_expert_actions.argtypes = [C.c_void_p, C.POINTER(MMState), C.c_void_p, C.c_void_p, C.c_void_p]
    L.mm_set_schedule.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mm_set_cycle_buffer.argtypes = [C.c_void_p, C.c_void_p]
    L.mm_set_placement_yaw.argtypes = [C.c_void_p, C.c_void_p]
    L.mm_sample_yaw.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mm_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    L.mm_sample_episode.argtypes = ([C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p] + [C.c_double] * 5 +
                                    [C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p] * 7)
    L.mm_host_staging.argtypes = [C.c_void_p, C.POINTER(MMStepOut)]
    L.mm_stage_timing.argtypes = [C.c_void_p, C.c_int32]
    L.mm_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mm_post_step.argtypes = [C.c_void_p, C.POINTER(MMState), C.POINTER(MMStepOut), C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int32, C.c_void_p]
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().mm_last_error().decode()}")
