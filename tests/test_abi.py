"""The C-ABI library loads and exports every symbol include/mm_manip.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from hostlib import REPO


def _declared():
    src = open(os.path.join(REPO, "include", "mm_manip.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mujoco_manip_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    for n in names:
        assert hasattr(L, n), n


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mujoco_manip_b200 import _lib

    L = _lib.lib()
    cfg = _lib.MMConfig(4, 0, 0, 32, 0, 500)
    h = C.c_void_p()
    assert L.mm_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in L.mm_last_error() or b"CPU" in L.mm_last_error()
    with pytest.raises((RuntimeError, ValueError)):
        from mujoco_manip_b200 import PickPlaceVecEnv

        PickPlaceVecEnv(4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "mujoco_manip_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp")) and "mm_emul" not in f:
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt, f
