"""ORACLE - TEST INFRASTRUCTURE ONLY.  CPU statement (numpy integer arithmetic) of the counter-based
placement sampler of mujoco_manip_b200/csrc/mm_rng.h.

Philox4x32-10 follows the published Random123 algorithm (Salmon et al., SC'11); its known-answer
vectors are checked in tests/test_rng.py.  The sampling rule above it restates the reference's
rejection sampler (mujoco_manip/randomization.py:70-98) and the 53-bit uniform construction numpy
uses for `Generator.uniform` ((a >> 5) * 2**26 + (b >> 6)) / 2**53, lo + (hi - lo) * u.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(ctr, key, rounds=10):
    c = [int(x) & MASK for x in ctr]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(rounds):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def u53(hi, lo):
    return float(((hi >> 5) << 26) | (lo >> 6)) / 9007199254740992.0


def place(seed, gid, episode, x_range=(-0.20, 0.20), y_range=(0.30, 0.45), min_sep=0.08, max_attempts=1000):
    """-> (xy[3,2], attempts) ; attempts = 0 when every attempt was rejected."""
    key = (seed & MASK, (seed >> 32) & MASK)
    sx, sy = np.float64(x_range[1]) - np.float64(x_range[0]), np.float64(y_range[1]) - np.float64(y_range[0])
    ms2 = np.float64(min_sep) * np.float64(min_sep)
    for a in range(max_attempts):
        w = []
        for k in range(3):
            w += philox4x32((gid & MASK, (gid >> 32) & MASK, episode & MASK, 4 * a + k), key)
        x = [np.float64(x_range[0]) + sx * np.float64(u53(w[2 * j], w[2 * j + 1])) for j in range(3)]
        y = [np.float64(y_range[0]) + sy * np.float64(u53(w[6 + 2 * j], w[7 + 2 * j])) for j in range(3)]
        ok = True
        for i in range(3):
            for j in range(i + 1, 3):
                dx, dy = x[i] - x[j], y[i] - y[j]
                if dx * dx + dy * dy < ms2:
                    ok = False
        if ok or a == max_attempts - 1:
            return np.array([[x[j], y[j]] for j in range(3)]), (a + 1 if ok else 0)
    return None, 0


def task_draw(seed, gid, episode, npool):
    key = (seed & MASK, (seed >> 32) & MASK)
    w = philox4x32((gid & MASK, (gid >> 32) & MASK, episode & MASK, 3), key)
    return (w[0] * npool) >> 32


def yaw(seed, gid, episode, o):
    """theta of cube o: uniform(0, 2 pi) from words 0,1 of block 4(o+1)+3 (randomization.py:55-62)."""
    key = (seed & MASK, (seed >> 32) & MASK)
    w = philox4x32((gid & MASK, (gid >> 32) & MASK, episode & MASK, 4 * (o + 1) + 3), key)
    return np.float64(0.0) + np.float64(6.283185307179586) * np.float64(u53(w[0], w[1]))
