"""Random streams: Philox known answers (Random123), the CPU statement of the placement sampler, and
the host sampler's bit-exactness with the reference's reset draws (golden)."""
import os

import numpy as np

from hostlib import GOLDEN
from mujoco_manip_b200.constants import TASK_SETS, task_indices
from mujoco_manip_b200.randomization import all_separated, sample_separated_positions
from oracle import philox


def test_philox_known_answers():
    assert philox.philox4x32((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    m = 0xFFFFFFFF
    assert philox.philox4x32((m, m, m, m), (m, m)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox.philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_placement_properties():
    seen = set()
    for gid in range(200):
        xy, att = philox.place(42, gid, 0)
        assert att >= 1
        assert all_separated([tuple(p) for p in xy], 0.08)
        assert (-0.2 <= xy[:, 0]).all() and (xy[:, 0] <= 0.2).all() and (0.30 <= xy[:, 1]).all() and (xy[:, 1] <= 0.45).all()
        seen.add(philox.task_draw(42, gid, 0, 9))
        assert np.array_equal(philox.place(42, gid, 0)[0], xy)          # counter based: reproducible
        assert not np.array_equal(philox.place(42, gid, 1)[0], xy)      # next episode differs
    assert seen == set(range(9))


def test_host_sampler_matches_reference_reset_draws():
    g = np.load(os.path.join(GOLDEN, "reset_seeds.npz"))
    for i, seed in enumerate(g["seeds"]):
        rng = np.random.default_rng(int(seed))
        xy = np.asarray(sample_separated_positions(rng, 3, (-0.20, 0.20), (0.30, 0.45)))
        np.testing.assert_array_equal(xy, g["obj_xy"][i])           # placement draws first ...
        t = TASK_SETS["all"][int(rng.integers(9))]                    # ... then the task draw
        assert tuple(task_indices(t)) == tuple(g["task"][i])
    np.testing.assert_array_equal([task_indices(t) for t in TASK_SETS["all"]], g["task_sets_all"])
    np.testing.assert_array_equal([task_indices(t) for t in TASK_SETS["cross"]], g["task_sets_cross"])


def test_sampler_failure_raises():
    import pytest

    with pytest.raises(RuntimeError):
        sample_separated_positions(np.random.default_rng(0), 3, (0.0, 0.01), (0.0, 0.01), 0.08)


def test_philox_yaw_stream():
    """randomize_yaw draws: words 0,1 of block 4(o+1)+3 of the (seed, env, episode) stream, theta = 2 pi u
    (values pinned from the first implementation; the device sampler is compared bit-exactly in the gpu tests)."""
    from oracle import philox

    got = [float(philox.yaw(42, 1000, 0, o)) for o in range(3)]
    assert got == [0.5236206822210825, 1.3999901980488791, 3.2421176688892728]
    assert float(philox.yaw(7, 2 ** 33 + 5, 3, 1)) == 1.0779784660538425
    # the yaw blocks are not the blocks of the first placement attempt or the task draw
    w = philox.philox4x32((1000, 0, 0, 7), (42, 0))
    assert got[0] == 6.283185307179586 * philox.u53(w[0], w[1])
    th = np.array([philox.yaw(1, e, 0, o) for e in range(200) for o in range(3)])
    assert th.min() >= 0 and th.max() < 2 * np.pi and abs(th.mean() - np.pi) < 0.3
