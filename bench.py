#!/usr/bin/env python3
"""Benchmark of the step hot path (BASELINE.json metric: env-steps/s, physics + IK + FSM).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]
                    [--precision f64|f32] [--group 8|16|32]

One "step" = one PickPlaceGymEnv.step (decode + 16 x (IK + physics substep) + forward + reward +
state observation, gym_env.py:536-581) for EVERY env of the batch = one mm_step launch.
Default workload = BASELINE.json configs[1]: 4096 envs per GPU, fixed task (obj_red, bin_red),
ee_pos_quat_g_rel random actions (SURVEY 8d distribution), state observations, 500-step episodes
with auto-reset.  Envs shard over GPUs with no data-path collective (weak scaling).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle restatement of the
reference path on the host cores (MuJoCo itself is not installable in this image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "env-steps/sec (physics+IK+FSM)"
UNIT = "env-steps/s"
MODE = "ee_pos_quat_g_rel"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--group", type=int, default=32)
    ap.add_argument("--workload", default="random", choices=["random", "fsm", "mixed"],
                    help="random = BASELINE configs[1] (headline); fsm = configs[2] scripted-expert dataset generation (extra); "
                         "mixed = configs[4] abs_pos + ee_pos_rot6d_g halves, tasks=cross, expert-driven, per-phase report (extra)")
    ap.add_argument("--mode", default=MODE, choices=["ee_pos_quat_g_rel", "ee_pos_rot6d_g_rel"],
                    help="action mode of the random workload (configs[3] uses ee_pos_rot6d_g_rel with --randomize)")
    ap.add_argument("--randomize", action="store_true", help="random workload: randomize_objects=True (device Philox)")
    ap.add_argument("--burnin", type=int, default=40, help="untimed steps before the warm-up: episodes are de-phased (staggered "
                    "step counters) and the contact load reaches its steady state (it climbs over the first ~20 steps after a reset)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(envs, mode=MODE, randomize=False, burnin=0):
    tag = "configs[1]" if (mode == MODE and not randomize) else ("configs[3]" if envs == 65536 else "configs[3]-style")
    return {"workload": f"{tag}: {envs} envs/GPU, task (obj_red,bin_red), {mode} random actions, state obs, "
                        f"500-step episodes, auto-reset{', randomized objects (Philox seed 1234)' if randomize else ''}",
            "steady_state": f"episode phases staggered uniformly over 0..499, {burnin} untimed burn-in steps before the warm-up",
            "envs_per_gpu": envs, "action_mode": mode,
            "substeps_per_env_step": 16, "l2": "flushed between timed steps (256 MiB write)"}


# ---- algorithmic FLOP model (DESIGN.md "FLOP model"; counters come from the oracle on the same workload) ----
def flops_per_env_step(stats):
    f = max(1, stats["forwards"])
    per = {k: stats[k] / f for k in ("ncon", "nefc", "newton_iters", "ls_evals", "narrow_tests", "ccd_tests", "nnzJ", "nnzJ2")}
    fixed = 14000.0                      # FK, CRB, RNE, actuation, M^-1, implicitfast, IK
    coll = 12.0 * 780 + 1200.0 * per["narrow_tests"] + 3000.0 * per["ccd_tests"] + 400.0 * per["ncon"]
    build = 600.0 * per["ncon"] + 40.0 * per["nefc"]
    newton = per["newton_iters"] * (4.0 * per["nnzJ"] + per["nnzJ2"] + 27 ** 3 / 3.0 + 2 * 27 ** 2) + 12.0 * per["nefc"] * per["ls_evals"]
    per_forward = fixed + coll + build + newton
    forwards_per_env_step = stats["forwards"] / max(1, stats["env_steps"])
    return per_forward * forwards_per_env_step, per


def cpu_sample(nthreads, seconds_target=12.0, mode=MODE):
    """Oracle restatement of the reference path on the host cores, bounded sample."""
    from oracle import oracle

    oracle.build()
    # calibrate: ~350 env-steps/s/core
    n_envs = max(nthreads, 8)
    n_steps = max(10, int(seconds_target * 300 * nthreads / n_envs))
    n_steps = min(n_steps, 500)
    v, st = oracle.bench_random_stats(n_envs, n_steps, mode=mode, seed=1234, nthreads=nthreads, flags=0)
    st["env_steps"] = n_envs * n_steps
    return v, st, f"{n_envs} envs x {n_steps} env-steps of configs[1] (same action distribution), {nthreads} threads"


class ClockSampler(threading.Thread):
    def __init__(self, idx):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.stop_flag, self.max_mhz = idx, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                bits = int(out[2].strip(), 16) if len(out) > 2 and out[2].strip().startswith("0x") else 0
                names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                         0x80: "hw_power_brake_slowdown"}
                for b, nme in names.items():
                    if bits & b:
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def reference_sample_mujoco(n_envs, n_steps, mode, seed):
    """The UNMODIFIED reference on real MuJoCo, when both are importable (they are not in the build image): one process
    per host core, each stepping PickPlaceGymEnv's hot path (decode -> 16 x (IK, ctrl, mj_step) -> mj_forward -> reward)
    without the renderer, as tests/test_pick_and_place.py:20-27 builds it.  Returns env-steps/s or None."""
    ref = os.path.join(REPO, "baseline", "_ref")  # where an installed copy of the reference would live
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import gymnasium  # noqa: F401
        import mujoco  # noqa: F401
        from mujoco_manip.gym_env import PickPlaceGymEnv  # noqa: F401
    except Exception:
        return None
    import multiprocessing as mp

    def work(q, k):
        import numpy as np
        from mujoco_manip.gym_env import PickPlaceGymEnv

        env = PickPlaceGymEnv(task=("obj_red", "bin_red"), action_mode=mode)
        env._get_obs = lambda: {}  # state-only path: no renders
        rng = np.random.default_rng(seed + k)
        env.reset(seed=seed + k)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            a = env.action_space.sample().astype(np.float32)
            a[:3] = rng.uniform(-0.3, 0.3, size=3)
            env.step(a)
        q.put(time.perf_counter() - t0)

    q = mp.Queue()
    ps = [mp.Process(target=work, args=(q, k)) for k in range(n_envs)]
    t0 = time.perf_counter()
    for p_ in ps:
        p_.start()
    for p_ in ps:
        p_.join()
    return n_envs * n_steps / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path on all host cores.  The real reference (MuJoCo +
    gymnasium) is tried first; in this image neither is installable, so the FP64 oracle restatement is timed instead
    (`kind: "port"`).  One bench step = one bounded sample of the workload; every step is timed on its own."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))  # seconds of CPU work per step
    kind, sample, times, counts = "port", "", [], []
    t_all = time.time()
    for i in range(args.warmup + args.steps):
        n_envs = max(cores, 8)
        n_steps = min(500, max(10, int(budget * 300 * cores / n_envs)))
        t0 = time.perf_counter()
        v = reference_sample_mujoco(cores, n_steps, args.mode, 1234 + i)
        if v is not None:
            kind, n_envs = "reference", cores
            sample = f"{cores} processes x {n_steps} env-steps of PickPlaceGymEnv (MuJoCo), renderer off"
        else:
            from oracle import oracle

            oracle.build()
            oracle.bench_random_stats(n_envs, n_steps, mode=args.mode, seed=1234 + i, nthreads=cores, flags=0)
            sample = (f"{n_envs} envs x {n_steps} env-steps of configs[1] (same action distribution), {cores} threads; MuJoCo is "
                      "not installable here, so this is the FP64 oracle restatement of the reference path")
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            counts.append(n_envs * n_steps)
    v = sum(counts) / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args.envs, args.mode, args.randomize),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + " per bench step"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "env_steps_per_bench_step": counts[0], "gpu_launches": 0, "wall_s": time.time() - t_all}
    print(json.dumps(line), flush=True)


def run_fsm(args, rank, world, local):
    """BASELINE.json configs[2] (extra line, not the headline): scripted-FSM expert rollouts, tasks = all (cycled by
    global env id), randomized objects (device Philox), abs_pos actions, finished episodes are reset in place."""
    import torch
    import torch.distributed as dist

    from mujoco_manip_b200 import PickPlaceVecEnv

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs
    env = PickPlaceVecEnv(n, device=dev, tasks="all", action_mode="abs_pos", randomize_objects=True, rng="philox", seed=42,
                          env_id_offset=rank * n, task_assignment="cycle", auto_reset=False, max_episode_steps=2000,
                          precision=args.precision, group=args.group)
    env.reset()
    episodes = torch.zeros(2, dtype=torch.float64, device=dev)

    def one():
        a = env.fsm_plan(16)
        obs, r, te, tr, info = env.step(a)
        done = env.fsm_state == 11
        episodes[0] += done.sum()
        episodes[1] += (done & info["success"]).sum()
        env.reset(mask=done.to(torch.uint8))

    for _ in range(args.warmup):
        one()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(episodes)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"metric": METRIC, "value": world * n * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                          "config": {"workload": f"configs[2]: {n} envs/GPU, tasks=all, randomize_objects seed 42, scripted-FSM "
                                                 "expert (plan(16) + step), state obs, episodes reset at FSM DONE"},
                          "episodes_finished": float(episodes[0]), "success_rate": float(episodes[1] / episodes[0].clamp(min=1))}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


PHASES = ("idle", "approaching", "grasping", "lifting", "transporting", "placing", "retreating", "done")
# FSM state index (1..11) -> phase index, as _STATE_TO_PHASE of pick_and_place.py:46-58
STATE_PHASE = (0, 0, 1, 2, 2, 3, 4, 4, 5, 5, 6, 7)


def run_mixed(args, rank, world, local):
    """BASELINE.json configs[4] (extra line): half the envs take abs_pos actions, half ee_pos_rot6d_g (absolute EE pose),
    tasks = cross (cycled by global env id), randomized objects, scripted expert so that grasp / lift contact phases
    dominate.  Per-phase figures come from the per-env busy cycles the step kernel writes for its scheduler."""
    import torch
    import torch.distributed as dist

    from mujoco_manip_b200 import PickPlaceVecEnv
    from mujoco_manip_b200.features import expert_action_encodings

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs
    half = n // 2
    envs = []
    for k, mode in enumerate(("abs_pos", "ee_pos_rot6d_g")):
        e = PickPlaceVecEnv(half, device=dev, tasks="cross", action_mode=mode, randomize_objects=True, rng="philox", seed=42,
                            env_id_offset=rank * n + k * half, task_assignment="cycle", auto_reset=False,
                            max_episode_steps=2000, precision=args.precision, group=args.group)
        e.reset()
        envs.append(e)
    sp = torch.tensor(STATE_PHASE, device=dev)
    steps_by_phase = torch.zeros(len(PHASES), dtype=torch.float64, device=dev)
    work_by_phase = torch.zeros(len(PHASES), dtype=torch.float64, device=dev)
    episodes = torch.zeros(2, dtype=torch.float64, device=dev)

    def one(count):
        for e in envs:
            a = e.fsm_plan(16)
            ph = sp[e.fsm_state.to(torch.int64)]
            if e.action_mode != "abs_pos":  # the expert's action in the absolute EE-pose encoding (generate_dataset.py:56-80)
                a = expert_action_encodings(e, a)[:, 8:18]
            _, _, _, _, info = e.step(a)
            done = e.fsm_state == 11
            if count:
                steps_by_phase.index_add_(0, ph, torch.ones(half, dtype=torch.float64, device=dev))
                work_by_phase.index_add_(0, ph, e._work.to(torch.float64))
                episodes[0] += done.sum()
                episodes[1] += (done & info["success"]).sum()
            e.reset(mask=done.to(torch.uint8))

    for _ in range(args.warmup):
        one(False)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one(True)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for x in (episodes, steps_by_phase, work_by_phase):
            dist.all_reduce(x)
    if rank == 0:
        ms = float(t[0])
        value = world * 2 * half * args.steps / (ms * 1e-3)
        share = (work_by_phase / work_by_phase.sum().clamp(min=1)).tolist()
        cnt = steps_by_phase.tolist()
        per_phase = {p: {"env_steps": int(c), "busy_share": round(s, 4),
                         # env-steps/s the job would reach if every env were in this phase (busy-cycle weighted)
                         "env_steps_per_s": (c / (s * ms * 1e-3)) if s > 0 else None}
                     for p, c, s in zip(PHASES, cnt, share)}
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                          "config": {"workload": f"configs[4]: {n} envs/GPU, half abs_pos + half ee_pos_rot6d_g, tasks=cross, "
                                                 "randomize_objects seed 42, scripted-FSM expert, episodes reset at FSM DONE"},
                          "episodes_finished": float(episodes[0]),
                          "success_rate": float(episodes[1] / episodes[0].clamp(min=1)), "per_phase": per_phase}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "fsm":
        run_fsm(args, rank, world, local)
        return
    if args.workload == "mixed":
        run_mixed(args, rank, world, local)
        return
    if not __import__("torch").cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    res = measure_random(args, rank, world, local, dev, args.envs, args.mode, args.randomize, args.steps, args.warmup,
                         with_e2e=not args.no_e2e, with_flops=True)
    extra = None
    if world > 1:  # BASELINE.json configs[3], the multi-GPU config, rides on the same line (nested object)
        extra = measure_random(args, rank, world, local, dev, 65536, "ee_pos_rot6d_g_rel", True, max(20, min(args.steps, 30)), 3,
                               with_e2e=False, with_flops=False)
    if rank == 0:
        line = finish_line(args, res, world, local)
        if extra is not None:
            line["configs3"] = {k: extra[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "overflow_envs",
                                                       "nonfinite_resets", "stage_ms_per_step", "episodes_finished")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_random(args, rank, world, local, dev, n, mode, randomize, steps, warmup, with_e2e, with_flops):
    """Random-action rollouts (configs[1] / configs[3] distribution) of n envs per GPU at steady state: staggered episode
    phases, `--burnin` untimed steps, then `steps` timed steps (CUDA events, L2 flushed between them)."""
    import torch
    import torch.distributed as dist

    from mujoco_manip_b200 import PickPlaceVecEnv, _lib

    env = PickPlaceVecEnv(n, device=dev, task=("obj_red", "bin_red"), action_mode=mode, reward_type="dense",
                          max_episode_steps=500, seed=1234, rng="philox", env_id_offset=rank * n, precision=args.precision,
                          group=args.group, auto_reset=True, randomize_objects=randomize)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # de-phase the episodes: without this every env truncates at the same step and the timed window sees one phase only
    env.state["step_count"].copy_(torch.randint(0, 500, (n, 1), device=dev, generator=gen, dtype=torch.int32))
    lo = torch.tensor([-0.3, 0.30, 0.30], device=dev, dtype=torch.float64)
    hi = torch.tensor([0.3, 0.65, 0.60], device=dev, dtype=torch.float64)
    T0 = env.state["tinit"][0]
    p0, R0 = T0[:3], T0[3:].reshape(3, 3)

    def make_actions():
        # world target uniform in a box above the table, expressed in the initial-EE frame; random
        # rotation part (decoded and ignored by the path); Bernoulli gripper
        w = lo + (hi - lo) * torch.rand((n, 3), device=dev, dtype=torch.float64, generator=gen)
        a = torch.zeros((n, _lib.ACTION_STRIDE), device=dev, dtype=torch.float32)
        a[:, :3] = ((w - p0) @ R0).float()
        if mode.startswith("ee_pos_rot6d"):
            a[:, 3:9] = torch.randn((n, 6), device=dev, generator=gen)
            a[:, 9] = (torch.rand(n, device=dev, generator=gen) > 0.5).float()
        else:
            q = torch.randn((n, 4), device=dev, generator=gen)
            a[:, 3:7] = q / q.norm(dim=1, keepdim=True)
            a[:, 7] = (torch.rand(n, device=dev, generator=gen) > 0.5).float()
        return a

    pool = [make_actions() for _ in range(16)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    mode_i = _lib.ACTION_MODES.index(mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(a):
        _lib.check(env._L.mm_step(env._h, C.byref(env._st), a.data_ptr(), mode_i, C.byref(env._out), env._stream()), "mm_step")
        env._post_step_autoreset()

    for i in range(args.burnin + warmup):
        one_step(pool[i % len(pool)])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stats0 = env.stats.clone()
    l0 = env.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    wall0 = time.time()
    for i in range(steps):
        flush.fill_(i & 0xFF)  # evict L2 (126 MB) between timed steps; outside the timed bracket of the step
        a = pool[(warmup + i) % len(pool)]
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        _lib.check(env._L.mm_step(env._h, C.byref(env._st), a.data_ptr(), mode_i, C.byref(env._out), env._stream()), "mm_step")
        kev[i][1].record(stream)
        env._post_step_autoreset()  # statistics, reset mask, Philox draw and reset of finished envs: library kernels
        ev[i][1].record(stream)
    barrier()
    wall = time.time() - wall0
    launches = env.launch_count() - l0
    stats1 = env.stats.clone()
    # per-stage device times of the overlapped plan: the same steps again with the library's per-launch events switched on
    # (outside the timed region: two event records per launch are not free)
    _lib.check(env._L.mm_stage_timing(env._h, 1), "mm_stage_timing")
    for i in range(steps):
        one_step(pool[(warmup + steps + i) % len(pool)])
    barrier()
    stage_ms = (C.c_double * 4)()
    stage_n = (C.c_longlong * 4)()
    _lib.check(env._L.mm_stage_times(env._h, stage_ms, stage_n), "mm_stage_times")
    _lib.check(env._L.mm_stage_timing(env._h, 0), "mm_stage_timing")
    step_ms = sum(a.elapsed_time(b) for a, b in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev)
    dstats = (stats1 - stats0).tolist()
    overflow_now = int((env.state["diag"][:, 2] != 0).sum())
    t = torch.tensor([step_ms, kern_ms] + list(stage_ms), device=dev, dtype=torch.float64)
    cnt = torch.tensor([dstats[0], dstats[4], dstats[5] + overflow_now, dstats[6]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    step_ms, kern_ms = float(t[0]), float(t[1])
    res = {"value": world * n * steps / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms / steps, "steps": steps, "warmup": warmup,
           "config": workload(n, mode, randomize, args.burnin), "kernel_ms_per_step": kern_ms / steps,
           "stage_ms_per_step": {"stage_a": float(t[2]) / steps, "convex": float(t[3]) / steps, "stage_c": float(t[4]) / steps,
                                 "stage_c_cta_per_env": float(t[5]) / steps},
           "stage_launches_per_step": {"stage_a": stage_n[0] / steps, "convex": stage_n[1] / steps, "stage_c": stage_n[2] / steps},
           "episodes_finished": float(cnt[0]), "nonfinite_resets": int(cnt[1]), "overflow_envs": int(cnt[2]),
           "failed_placements": int(cnt[3]), "gpu_launches": launches, "wall": wall, "n": n, "mode": mode}

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D actions, step, D2H obs / reward / flags ----
    if with_e2e:
        h_act = [p.cpu().pin_memory() for p in pool[:8]]
        h_obs = torch.zeros((n, _lib.OBS_DIM), dtype=torch.float32).pin_memory()
        h_rew = torch.zeros(n, dtype=torch.float32).pin_memory()
        h_fl = torch.zeros((3, n), dtype=torch.uint8).pin_memory()
        ksteps = max(3, steps // 3)

        def host_step(i):
            env.step_host(h_act[i % len(h_act)], h_obs, h_rew, h_fl)  # mm_step_host + device-side restart of finished envs

        for i in range(3):
            host_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(ksteps):
            host_step(i)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": world * n * ksteps / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": n * _lib.ACTION_STRIDE * 4,
                      "d2h_bytes_per_step": n * (_lib.OBS_DIM * 4 + 4 + 3), "steps": ksteps}
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        res["clocks"] = sampler.result()

    # ---- algorithmic FLOPs: the kernel source itself, instantiated with an operation-counting scalar (tests/mm_emul.cpp,
    # host build, one lane), stepped once from the states a sample of the benchmarked envs is in right now ----
    if with_flops and rank == 0:
        try:
            res["flops"] = count_flops(env, pool[(warmup + steps) % len(pool)], mode)
        except Exception as exc:  # the counting build is measurement infrastructure; the bench line survives without it
            res["flops"] = {"error": str(exc)[:200]}
        # per-kernel durations for the roofline: in the timed region the chunks' stage launches overlap on several
        # streams, so their event times include waiting for SMs; here the same batch, continued from the very same
        # states, takes three more steps on ONE stream and ONE chunk - every stage kernel timed alone, back to back
        old = os.environ.get("MM_STREAMS"), os.environ.get("MM_CHUNK")
        os.environ["MM_STREAMS"], os.environ["MM_CHUNK"] = "1", str(n)
        try:
            solo = PickPlaceVecEnv(n, device=dev, task=("obj_red", "bin_red"), action_mode=mode, reward_type="dense",
                                   max_episode_steps=500, seed=1234, rng="philox", env_id_offset=rank * n,
                                   precision=args.precision, group=args.group, auto_reset=True, randomize_objects=randomize)
            solo.reset()
            solo.set_state(env.get_state())
            solo._task.copy_(env._task)
            torch.cuda.synchronize(dev)
            _lib.check(solo._L.mm_stage_timing(solo._h, 1), "mm_stage_timing")
            for i in range(3):
                flush.fill_(i)
                a = pool[(warmup + steps + i) % len(pool)]
                _lib.check(solo._L.mm_step(solo._h, C.byref(solo._st), a.data_ptr(), mode_i, C.byref(solo._out), solo._stream()), "mm_step")
                solo._post_step_autoreset()
            sm, sn = (C.c_double * 4)(), (C.c_longlong * 4)()
            _lib.check(solo._L.mm_stage_times(solo._h, sm, sn), "mm_stage_times")
            res["solo"] = {"stage_a_ms_per_launch": sm[0] / max(1, sn[0]), "convex_ms_per_launch": sm[1] / max(1, sn[1]),
                           "stage_c_ms_per_launch": sm[2] / max(1, sn[2]), "launches_per_step": sn[2] / 3.0,
                           "step_ms": (sm[0] + sm[1] + sm[2]) / 3.0}
            solo.close()
        except Exception as exc:
            res["solo"] = {"error": str(exc)[:200]}
        finally:
            for k, v in zip(("MM_STREAMS", "MM_CHUNK"), old):
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    res["env_handle"] = env
    return res


def count_flops(env, actions, mode, sample=128):
    import numpy as np

    sys.path.insert(0, os.path.join(REPO, "tests"))
    import hostlib

    n = env.num_envs
    idx = np.linspace(0, n - 1, sample).astype(np.int64)
    em = hostlib.EmulEnv(sample, mode=mode, reward="dense", max_steps=500)
    for k, v in env.state.items():
        if k in em.st:
            em.st[k][...] = v[idx].cpu().numpy().reshape(em.st[k].shape)
    em.tgt[...] = 0
    ncon = float(env.state["diag"][:, 0].double().mean())
    ncon_sample = float(env.state["diag"][idx, 0].double().mean())
    fl = em.step_counted(actions[idx].cpu().numpy())
    per = fl.astype(np.float64) / sample
    return {"per_env_step": float(per.sum()), "stage_a": float(per[0]), "convex": float(per[1]), "stage_c": float(per[2]),
            "sample_envs": sample, "mean_contacts_in_batch": ncon, "mean_contacts_in_sample": ncon_sample,
            "how": "operation-counting scalar through the kernel source (tests/mm_emul.cpp, 1 lane): adds, multiplies, divides, "
                   "square roots and trigonometric calls count 1 each (FMA = 2)"}


def finish_line(args, res, world, local):
    from mujoco_manip_b200 import _lib

    env = res.pop("env_handle")
    n, mode = res.pop("n"), res.pop("mode")
    fp64 = args.precision == "f64"
    peak = C.c_double()
    _lib.check(env._L.mm_measure_fma_peak(local, 1 if fp64 else 0, C.byref(peak)), "mm_measure_fma_peak")
    cpu = None
    model_flops = per = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, st, sample = cpu_sample(cores, mode=mode)
        model_flops, per = flops_per_env_step(st)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": sample + "; FP64 oracle restatement of the reference path (MuJoCo not installable here)"}
    fl = res.pop("flops", None) or {}
    solo = res.pop("solo", None) or {}
    stage = res["stage_ms_per_step"]
    launches_c = max(1.0, solo.get("launches_per_step") or res["stage_launches_per_step"]["stage_c"])
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    hbm_peak = peaks.get("hbm_gbs", 6548.5)
    traffic = None
    tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("envs") == n and tj.get("precision") == args.precision:
            traffic = tj.get("stage_c_dram_bytes_per_launch")
    state_bytes = 8 * (30 + 27 + 8 + 27) * 2 + 8 * 12 * 2 + 4 * 10 + 4 * 85 + 4 + 3 + 4 * 8  # FP64 state in+out, tinit/eepose, action, obs, reward/flags, ints
    kern_s = res["kernel_ms_per_step"] * 1e-3
    c_flops = fl.get("stage_c")
    # duration of one stage-C launch timed ALONE (single stream, single chunk, same states); the overlapped timed region
    # gives only sums that exceed wall time
    c_launch_s = (solo["stage_c_ms_per_launch"] if "stage_c_ms_per_launch" in solo
                  else stage["stage_c"] / max(1.0, res["stage_launches_per_step"]["stage_c"])) * 1e-3
    roof = {"bound": "fp64-cuda-core" if fp64 else "fp32-cuda-core",
            "kernel": f"k_stage_c<{args.precision},G={args.group}> (contact assembly, constraint rows, Newton solver, integration): "
                      "the dominant of the three stage kernels; one launch = one of the 17 rounds of the whole batch, timed alone "
                      "(single stream) right after the timed region, from the same states",
            # algorithmic FLOPs of one launch (one of the 17 rounds of one chunk: stage-C FLOPs per env-step x envs / launches per
            # step) / its mean CUDA-event duration
            "achieved": (c_flops * n / launches_c / c_launch_s * 1e-12) if c_flops else None,
            "peak": peak.value, "unit": "TFLOP/s",
            "frac": (c_flops * n / launches_c / c_launch_s * 1e-12 / peak.value) if c_flops else None,
            "peak_source": "measured live: dependent-FMA microkernel mm_measure_fma_peak (MEASURED_PEAKS.json has no CUDA-core figure)",
            "kernel_ms_per_launch": c_launch_s * 1e3, "launches_per_step": launches_c, "kernels_timed_alone": solo or None,
            "flops_per_env_step": fl or None,
            "whole_step": {"achieved": (fl["per_env_step"] * n / kern_s * 1e-12) if fl.get("per_env_step") else None,
                           "frac": (fl["per_env_step"] * n / kern_s * 1e-12 / peak.value) if fl.get("per_env_step") else None,
                           "ms": kern_s * 1e3, "overlapped_stage_ms_sum": stage,
                           "note": "mm_step of the timed region: the chunks run on several streams, stage launches overlap"},
            "oracle_model": {"flops_per_env_step": model_flops, "per_forward_counters": per,
                             "note": "round-1 model with hand coefficients on the oracle's dense counters, kept for continuity"},
            "traffic": traffic,
            "hbm": {"achieved": state_bytes * n / kern_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": state_bytes * n / kern_s * 1e-9 / hbm_peak, "bytes_per_env_step": state_bytes}}
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"], "warmup": res["warmup"],
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic", "config": res["config"], "substeps_per_s": res["value"] * 16,
            "e2e": res.get("e2e"), "gpu_launches": res["gpu_launches"], "roofline": roof, "cpu_baseline": cpu,
            "clocks": res.get("clocks"), "overflow_envs": res["overflow_envs"], "nonfinite_resets": res["nonfinite_resets"],
            "failed_placements": res["failed_placements"], "episodes_finished": res["episodes_finished"],
            "wall_s_timed_region": res["wall"], "group": args.group}
    return line


if __name__ == "__main__":
    main()
