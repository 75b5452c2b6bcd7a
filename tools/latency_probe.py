#!/usr/bin/env python3
"""Per-env cycles spent in the stage kernels of a step (GPU box): which envs are the long poles, and in which stage?
(mm_set_cycle_buffer: [total, stage A, convex stage (summed over the warps that ran the env's pairs), stage C].)
    python tools/latency_probe.py [--envs 4096] [--steps 40] [--group 16] [--precision f64]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujoco_manip_b200 import PickPlaceVecEnv, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--group", type=int, default=32)
ap.add_argument("--precision", default="f64")
a = ap.parse_args()
dev = torch.device("cuda:0")
env = PickPlaceVecEnv(a.envs, device=dev, task=("obj_red", "bin_red"), precision=a.precision, group=a.group, seed=1234)
env.reset()
cyc9 = torch.zeros((a.envs, 9), dtype=torch.int64, device=dev)
cyc = cyc9[:, 0]
_lib.check(env._L.mm_set_cycle_buffer(env._h, cyc9.data_ptr()), "cycles")
gen = torch.Generator(device=dev).manual_seed(1234)
T0 = env.state["tinit"][0]
p0, R0 = T0[:3], T0[3:].reshape(3, 3)
lo = torch.tensor([-0.3, 0.30, 0.30], device=dev, dtype=torch.float64)
hi = torch.tensor([0.3, 0.65, 0.60], device=dev, dtype=torch.float64)
for t in range(a.steps):
    w = lo + (hi - lo) * torch.rand((a.envs, 3), device=dev, dtype=torch.float64, generator=gen)
    act = torch.zeros((a.envs, 10), device=dev)
    act[:, :3] = ((w - p0) @ R0).float()
    act[:, 6] = 1.0
    act[:, 7] = (torch.rand(a.envs, device=dev, generator=gen) > 0.5).float()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cyc9.zero_()
    e0.record()
    env.step(act)
    e1.record()
    torch.cuda.synchronize()
    if t % 5 == 4 or t == a.steps - 1:
        ph = cyc9[:, 1:4].double() / 1.965e6  # ms at 1965 MHz: stage A, convex, stage C
        c = ph.sum(dim=1)
        q = torch.quantile(c, torch.tensor([0.5, 0.9, 0.99, 1.0], device=dev, dtype=torch.float64))
        ncon = env.state["diag"][:, 0].double()
        worst = int(torch.argmax(c))
        print(f"step {t}: step {e0.elapsed_time(e1):.1f} ms | per-env busy ms p50 {q[0]:.2f} p90 {q[1]:.2f} p99 {q[2]:.2f} max {q[3]:.2f} "
              f"| sum/SMs {float(c.sum()) / 148:.1f} | ncon mean {float(ncon.mean()):.1f} max {int(ncon.max())} "
              f"| worst env ncon {int(ncon[worst])} iters {int(env.state['diag'][worst, 1])}")
        order = torch.argsort(c)
        med, slow = order[a.envs // 2 - 50: a.envs // 2 + 50], order[-a.envs // 20:]
        names = ["stage A", "convex", "stage C"]
        print("      stage ms   " + "  ".join(f"{n:>8s}" for n in names))
        print("      mean       " + "  ".join(f"{float(ph[:, k].mean()):8.2f}" for k in range(3)))
        print("      median env " + "  ".join(f"{float(ph[med, k].mean()):8.2f}" for k in range(3)))
        print("      slowest 5% " + "  ".join(f"{float(ph[slow, k].mean()):8.2f}" for k in range(3)))
# which envs carry the load: share of envs / of each stage's busy time per contact-count class (last step)
edges = [0, 8, 16, 24, 32, 48, 64, 96, 1 << 30]
print("contacts      envs     stage A ms (share)   convex ms (share)   stage C ms (share)")
for lo_, hi_ in zip(edges[:-1], edges[1:]):
    sel = (ncon >= lo_) & (ncon < hi_)
    if int(sel.sum()) == 0:
        continue
    cells = "   ".join(f"{float(ph[sel, k].mean()):6.3f} ({100 * float(ph[sel, k].sum() / ph[:, k].sum()):4.1f} %)" for k in range(3))
    print(f"{lo_:3d}..{min(hi_, 999):3d}  {100 * float(sel.double().mean()):6.1f} %   {cells}")
