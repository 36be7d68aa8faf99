"""MT19937 restatement of the reference's minibatch index stream (oracle).

The reference draws a minibatch with
``random.choices(range(K), cum_weights=np.arange(1, K+1), k=B)``
(src/data_set.py:31,67) after ``random.seed(seed)``
(src/pytorch_eigen_solver.py:61).  With those uniform cumulative weights the
bisection in CPython's ``random.choices`` reduces to

    idx = min(floor(u * K), K - 1),   u = ((a >> 5) * 2**26 + (b >> 6)) / 2**53

with a, b two consecutive 32-bit MT19937 outputs (CPython _randommodule.c
``random_random``), evaluated in IEEE double.  Seeding is CPython's
``init_by_array`` over the 32-bit little-endian words of ``abs(seed)``.

Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

N, M = 624, 397
_U32 = 0xFFFFFFFF


def seed_state(seed: int) -> np.ndarray:
    """MT state (624 uint32) after ``random.seed(seed)`` for an int seed; position = 624."""
    seed = abs(int(seed))
    key = []
    while True:
        key.append(seed & _U32)
        seed >>= 32
        if seed == 0:
            break
    mt = [0] * N
    mt[0] = 19650218
    for i in range(1, N):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & _U32
    i, j = 1, 0
    for _ in range(max(N, len(key))):
        mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & _U32
        i += 1; j += 1
        if i >= N:
            mt[0] = mt[N - 1]; i = 1
        if j >= len(key):
            j = 0
    for _ in range(N - 1):
        mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & _U32
        i += 1
        if i >= N:
            mt[0] = mt[N - 1]; i = 1
    mt[0] = 0x80000000
    return np.array(mt, dtype=np.uint32)


def _twist(mt: np.ndarray) -> np.ndarray:
    """Regenerate the 624-word block (scalar reference implementation)."""
    mt = [int(v) for v in mt]
    for kk in range(N):
        y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % N] & 0x7FFFFFFF)
        mt[kk] = mt[(kk + M) % N] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
    return np.array(mt, dtype=np.uint32)


def _temper(y: np.ndarray) -> np.ndarray:
    y = y.astype(np.uint64)
    y ^= (y >> 11)
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= (y >> 18)
    return (y & _U32).astype(np.uint32)


class MT19937:
    def __init__(self, seed=None, state=None, pos=N):
        self.mt = seed_state(seed) if state is None else np.array(state, dtype=np.uint32)
        self.pos = pos

    def words(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        filled = 0
        while filled < n:
            if self.pos >= N:
                self.mt = _twist(self.mt)
                self.pos = 0
            take = min(n - filled, N - self.pos)
            out[filled:filled + take] = _temper(self.mt[self.pos:self.pos + take])
            self.pos += take
            filled += take
        return out

    def random(self, n: int) -> np.ndarray:
        w = self.words(2 * n).astype(np.float64)
        a = np.floor(w[0::2] / 32.0)
        b = np.floor(w[1::2] / 64.0)
        return (a * 67108864.0 + b) / 9007199254740992.0

    def minibatch(self, K: int, B: int) -> np.ndarray:
        u = self.random(B)
        return np.minimum(np.floor(u * float(K)).astype(np.int64), K - 1)
