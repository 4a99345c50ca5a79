"""TEST INFRASTRUCTURE -- not part of the product path.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) restated in numpy so the tests can reproduce the kernels' counter-based streams on
the CPU bit for bit.  The reference repo has no RNG of its own on this path (gym-PBN draws
from python ``random``/numpy); the Philox stream is part of the *build's* contract
(BASELINE.json north_star), so this file restates the published algorithm, pinned by the
Random123 known-answer vectors in tests/test_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R.  Counter words are array-likes (broadcastable) of uint32
    values, key words are python ints.  Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def philox_scalar(ctr, key, rounds=10):
    """Single-counter convenience wrapper: ctr = 4 ints, key = 2 ints -> list of 4 ints."""
    out = philox4x32(*[np.array([c]) for c in ctr], key[0], key[1], rounds)
    return [int(x[0]) for x in out]
