#!/usr/bin/env python3
"""How often do the fixed-size per-env lists of the kernels overflow (GPU box)?  Random-action workload of bench.py.
    python tools/overflow_probe.py [--envs 16384] [--steps 500]
Bits of diag[:, 2]: 1 = broad-phase survivor list (MAXSURV = 384), 2 = contact list (MAXCON = 256), 4 = body-pair slots
(MAXPAIR = 96 = every (class, class) key of the model: cannot happen), 8 = convex-pair queue of the chunk full."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujoco_manip_b200 import PickPlaceVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--steps", type=int, default=500)
a = ap.parse_args()
dev = torch.device("cuda:0")
n = a.envs
env = PickPlaceVecEnv(n, device=dev, task=("obj_red", "bin_red"), seed=1234)
env.reset()
gen = torch.Generator(device=dev).manual_seed(1234)
T0 = env.state["tinit"][0]
p0, R0 = T0[:3], T0[3:].reshape(3, 3)
lo = torch.tensor([-0.3, 0.30, 0.30], device=dev, dtype=torch.float64)
hi = torch.tensor([0.3, 0.65, 0.60], device=dev, dtype=torch.float64)
ever = torch.zeros(n, dtype=torch.int32, device=dev)
hits = torch.zeros(4, dtype=torch.int64, device=dev)
prev = torch.zeros(n, dtype=torch.int32, device=dev)
mx = 0
for t in range(a.steps):
    w = lo + (hi - lo) * torch.rand((n, 3), device=dev, dtype=torch.float64, generator=gen)
    act = torch.zeros((n, 10), device=dev)
    act[:, :3] = ((w - p0) @ R0).float()
    act[:, 6] = 1
    act[:, 7] = (torch.rand(n, device=dev, generator=gen) > 0.5).float()
    env.step(act)
    d = env.state["diag"][:, 2].clone()  # sticky within an episode, cleared by the reset
    new = d & ~prev
    prev = d
    ever |= d
    for b in range(4):
        hits[b] += ((new >> b) & 1).sum()
    mx = max(mx, int(env.state["diag"][:, 0].max()))
tot = n * a.steps
print(json.dumps({"envs": n, "steps": a.steps, "env_steps": tot, "episodes_with_overflow": {
    "survivors(MAXSURV)": int(hits[0]), "contacts(MAXCON)": int(hits[1]), "body_pairs(MAXPAIR)": int(hits[2]),
    "convex_queue": int(hits[3])}, "stats_overflow_episodes": float(env.stats[5]),
    "body_pair_overflows_per_env_step": float(hits[2]) / tot, "envs_ever_overflowed": int((ever != 0).sum()), "max_ncon": mx,
    "episodes": float(env.stats[0]), "nonfinite_resets": float(env.stats[4])}))
