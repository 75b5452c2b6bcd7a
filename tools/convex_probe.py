#!/usr/bin/env python3
"""Where the convex narrow phase (GJK / EPA) of the slowest envs spends its time (GPU box).  Needs the instrumented build:
    tools/build_variant.sh prof -DMM_PROF_CONVEX
    MM_LIB_PATH=$PWD/mujoco_manip_b200/_C/variants/libmm_prof.so python tools/convex_probe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujoco_manip_b200 import PickPlaceVecEnv, _lib  # noqa: E402

n = 4096
dev = torch.device("cuda:0")
env = PickPlaceVecEnv(n, device=dev, task=("obj_red", "bin_red"), seed=1234)
env.reset()
cyc9 = torch.zeros((n, 9), dtype=torch.int64, device=dev)
_lib.check(env._L.mm_set_cycle_buffer(env._h, cyc9.data_ptr()), "cycles")
gen = torch.Generator(device=dev).manual_seed(1234)
T0 = env.state["tinit"][0]
p0, R0 = T0[:3], T0[3:].reshape(3, 3)
lo = torch.tensor([-0.3, 0.30, 0.30], device=dev, dtype=torch.float64)
hi = torch.tensor([0.3, 0.65, 0.60], device=dev, dtype=torch.float64)
for t in range(20):
    w = lo + (hi - lo) * torch.rand((n, 3), device=dev, dtype=torch.float64, generator=gen)
    act = torch.zeros((n, 10), device=dev)
    act[:, :3] = ((w - p0) @ R0).float()
    act[:, 6] = 1.0
    act[:, 7] = (torch.rand(n, device=dev, generator=gen) > 0.5).float()
    env.step(act)
torch.cuda.synchronize()
tot = cyc9[:, 0].double() / 1.965e6
ms = cyc9[:, 1:].double() / 1.965e6
cnt = cyc9[:, 1:].double() / 64.0  # counters are exported << 6 like the timers
order = torch.argsort(tot)
groups = {"all": order, "median 100": order[n // 2 - 50: n // 2 + 50], "slowest 5%": order[-n // 20:], "slowest 16": order[-16:]}
print("per env-step (17 forward passes): total ms | convex section ms = wait + shapes + gjk + epa (+ merge) | calls")
for name, idx in groups.items():
    print(f"{name:12s} total {float(tot[idx].mean()):6.2f} | section {float(ms[idx, 7].mean()):6.2f} = wait {float(ms[idx, 0].mean()):5.2f} "
          f"+ shapes {float(ms[idx, 1].mean()):5.2f} + gjk {float(ms[idx, 2].mean()):5.2f} + epa {float(ms[idx, 3].mean()):5.2f} | "
          f"gjk calls {float(cnt[idx, 4].mean()):6.1f} epa calls {float(cnt[idx, 5].mean()):6.1f} epa iters {float(cnt[idx, 6].mean()):7.1f}")
