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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(envs, mode=MODE, randomize=False):
    tag = "configs[1]" if (mode == MODE and not randomize) else "configs[3]-style"
    return {"workload": f"{tag}: {envs} envs/GPU, task (obj_red,bin_red), {mode} random actions, state obs, "
                        f"500-step episodes, auto-reset{', randomized objects (Philox seed 1234)' if randomize else ''}",
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.time()
    vals = []
    sample = ""
    for i in range(max(1, min(args.steps, 3)) + (1 if args.warmup else 0)):
        v, st, sample = cpu_sample(cores, seconds_target=8.0, mode=args.mode)
        if i or not args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * args.envs / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args.envs, args.mode, args.randomize),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample + " per timed sample; MuJoCo is not installable here, so this is the FP64 "
                                                "oracle restatement of the reference path"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
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
    MODE = args.mode

    import torch
    import torch.distributed as dist

    from mujoco_manip_b200 import PickPlaceVecEnv, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs
    env = PickPlaceVecEnv(n, device=dev, task=("obj_red", "bin_red"), action_mode=MODE, reward_type="dense",
                          max_episode_steps=500, seed=1234, rng="philox", env_id_offset=rank * n, precision=args.precision,
                          group=args.group, auto_reset=True, randomize_objects=args.randomize)
    env.reset()
    total = args.steps + args.warmup

    # synthetic actions: world target uniform in a box above the table, expressed in
    # the initial-EE frame; random unit quaternion (decoded and ignored by the path); Bernoulli gripper
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    lo = torch.tensor([-0.3, 0.30, 0.30], device=dev, dtype=torch.float64)
    hi = torch.tensor([0.3, 0.65, 0.60], device=dev, dtype=torch.float64)
    T0 = env.state["tinit"][0]
    p0, R0 = T0[:3], T0[3:].reshape(3, 3)

    def make_actions():
        w = lo + (hi - lo) * torch.rand((n, 3), device=dev, dtype=torch.float64, generator=gen)
        a = torch.zeros((n, _lib.ACTION_STRIDE), device=dev, dtype=torch.float32)
        a[:, :3] = ((w - p0) @ R0).float()  # R0^T (w - p0)
        if MODE.startswith("ee_pos_rot6d"):  # any two 3-vectors: rotmat_from_6d orthonormalises them
            a[:, 3:9] = torch.randn((n, 6), device=dev, generator=gen)
            a[:, 9] = (torch.rand(n, device=dev, generator=gen) > 0.5).float()
        else:
            q = torch.randn((n, 4), device=dev, generator=gen)
            a[:, 3:7] = q / q.norm(dim=1, keepdim=True)
            a[:, 7] = (torch.rand(n, device=dev, generator=gen) > 0.5).float()
        return a

    pool = [make_actions() for _ in range(min(total, 32))]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        env.step(pool[i % len(pool)])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = env.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 (126 MB) between timed steps; outside the timed bracket of the step
        a = pool[(args.warmup + i) % len(pool)]
        ev[i][0].record(stream)
        # dominant kernel timed on its own (same stream): the step launch of mm_step
        env._schedule()
        kev[i][0].record(stream)
        _lib.check(env._L.mm_step(env._h, C.byref(env._st), a.data_ptr(), _lib.ACTION_MODES.index(MODE), C.byref(env._out),
                                  env._stream()), "mm_step")
        kev[i][1].record(stream)
        env._post_step_autoreset()
        ev[i][1].record(stream)
    barrier()
    wall = time.time() - wall0
    launches = env.launch_count() - l0
    step_ms = sum(a.elapsed_time(b) for a, b in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev)
    t = torch.tensor([step_ms, kern_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(t[0]), float(t[1])
    value = world * n * args.steps / (step_ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D actions, step, D2H obs/reward/flags ----
    e2e = None
    if not args.no_e2e:
        h_act = [p.cpu().pin_memory() for p in pool[:8]]
        h_obs = torch.zeros((n, _lib.OBS_DIM), dtype=torch.float32).pin_memory()
        h_rew = torch.zeros(n, dtype=torch.float32).pin_memory()
        h_fl = torch.zeros((3, n), dtype=torch.uint8).pin_memory()
        ksteps = max(3, args.steps // 3)

        def host_step(i):
            env._schedule()
            _lib.check(env._L.mm_step_host(env._h, C.byref(env._st), h_act[i % len(h_act)].data_ptr(),
                                           _lib.ACTION_MODES.index(MODE), h_obs.data_ptr(), h_rew.data_ptr(),
                                           h_fl[0].data_ptr(), h_fl[1].data_ptr(), h_fl[2].data_ptr(), env._stream()),
                       "mm_step_host")
            done = (h_fl[0] | h_fl[1]).bool()
            if bool(done.any()):  # host-side auto-reset decision, as a user of the host API would make it
                env.reset(mask=done.to(dev).to(torch.uint8))

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
        e2e = {"value": world * n * ksteps / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": n * _lib.ACTION_STRIDE * 4,
               "d2h_bytes_per_step": n * (_lib.OBS_DIM * 4 + 4 + 3), "steps": ksteps}
    clocks = None
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        clocks = sampler.result()

    if rank == 0:
        fp64 = args.precision == "f64"
        peak = C.c_double()
        _lib.check(env._L.mm_measure_fma_peak(local, 1 if fp64 else 0, C.byref(peak)), "mm_measure_fma_peak")
        cpu = None
        flops = None
        per = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, st, sample = cpu_sample(cores, mode=MODE)
            flops, per = flops_per_env_step(st)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": sample + "; FP64 oracle restatement of the reference path (MuJoCo not installable here)"}
        kern_s = kern_ms * 1e-3 / args.steps
        # DRAM bytes of one launch of the step kernel from the committed `ncu --set full` capture (same config)
        traffic = None
        tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("envs") == n and tj.get("precision") == args.precision:
                traffic = tj.get("dram_bytes_per_launch")
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
        hbm_peak = peaks.get("hbm_gbs", 6548.5)
        state_bytes = 8 * (30 + 27 + 8 + 27) * 2 + 8 * 12 * 2 + 4 * 10 + 4 * 85 + 4 + 3 + 4 * 8  # FP64 state in+out, tinit/eepose, action, obs, reward/flags, ints
        roof = {"bound": "fp64-cuda-core" if fp64 else "fp32-cuda-core", "kernel": f"k_step<{args.precision},G={args.group}>",
                "achieved": (flops * n / kern_s * 1e-12) if flops else None, "peak": peak.value, "unit": "TFLOP/s",
                "frac": (flops * n / kern_s * 1e-12 / peak.value) if flops else None,
                "peak_source": "measured live: dependent-FMA microkernel mm_measure_fma_peak (MEASURED_PEAKS.json has no CUDA-core figure)",
                "flops_per_env_step": flops, "per_forward_counters": per, "kernel_ms_per_launch": kern_s * 1e3,
                "traffic": traffic,
                "hbm": {"achieved": state_bytes * n / kern_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": state_bytes * n / kern_s * 1e-9 / hbm_peak, "bytes_per_env_step": state_bytes}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic", "config": workload(n, MODE, args.randomize), "substeps_per_s": value * 16,
                "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "wall_s_timed_region": wall, "group": args.group}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
