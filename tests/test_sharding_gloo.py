"""World-size-2 gloo test of the only collective of the system (episode statistics) and of the env-id
sharding rule.  Envs need no exchange on the data path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mujoco_manip_b200.stats import gather_stats, shard_range, summarize


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    local = torch.tensor([hi - lo, rank + 1, 100.0 * (rank + 1), -2.0 * (rank + 1), 0.0, 0.0, 0.0, 0.0], dtype=torch.float64)
    allst = gather_stats(local)
    q.put((rank, lo, hi, allst.tolist(), summarize(allst)))
    dist.destroy_process_group()


def test_stats_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, a0, s0), (r1, lo1, hi1, a1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 5, 5, 10)
    assert a0 == a1 and len(a0) == 2
    assert s0["episodes"] == 10 and abs(s0["success_rate"] - 0.3) < 1e-12 and abs(s0["mean_length"] - 30.0) < 1e-12


def test_shard_range_covers_everything():
    for n in (1, 7, 4096, 65536 * 8 + 3):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_single_process_passthrough():
    s = torch.tensor([4.0, 2.0, 40.0, 8.0, 0.0, 1.0, 0.0, 0.0], dtype=torch.float64)
    out = gather_stats(s)
    assert out.shape == (1, 8) and summarize(out)["overflow_episodes"] == 1.0
    assert summarize(out)["success_rate"] == 0.5
