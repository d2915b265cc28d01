"""R's default random number stream as used by ``set.seed(1988); sample(nsnp, k)`` in
R/computePairwiseMI.R:95-96 (Mersenne-Twister, sample.kind = "Rejection"), so that the host side can
form ``lr_links_approx`` without an R process.  Restates base R's RNG.c / random.c; NOT verified against a
live R in this image (none is installed) -- pass ``lr_links_approx=`` explicitly to bypass it."""
from __future__ import annotations

import math

import numpy as np

_N, _M = 624, 397


class RMersenne:
    def __init__(self, seed: int):
        seed &= 0xFFFFFFFF
        for _ in range(50):  # Randomize(): initial scrambling
            seed = (69069 * seed + 1) & 0xFFFFFFFF
        st = []
        for _ in range(_N + 1):  # RNG_Init(): i_seed[0] is mti, then the 624 state words
            seed = (69069 * seed + 1) & 0xFFFFFFFF
            st.append(seed)
        self.mt = st[1:]
        self.mti = _N  # FixupSeeds(): dummy[0] = 624

    def _next_u32(self) -> int:
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N - _M):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + _M] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            for kk in range(_N - _M, _N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + (_M - _N)] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            y = (mt[_N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
            mt[_N - 1] = mt[_M - 1] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y

    def unif_rand(self) -> float:
        v = self._next_u32() * 2.3283064365386963e-10
        if v <= 0.0:
            return 0.5 * 2.328306437080797e-10
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return v

    def _rbits(self, bits: int) -> int:
        v, n = 0, 0
        while n <= bits:
            v = 65536 * v + int(math.floor(self.unif_rand() * 65536))
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn: int) -> int:
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample(self, n: int, k: int) -> np.ndarray:
        """sample(n, k) without replacement (n <= 1e7: partial Fisher-Yates); 1-based."""
        x = list(range(n))
        out = np.empty(k, dtype=np.int64)
        for i in range(k):
            j = self.unif_index(n)
            out[i] = x[j] + 1
            n -= 1
            x[j] = x[n]
        return out


def lr_links_approx(POS: np.ndarray, g: float, sr_dist: float, seed: int = 1988) -> float:
    """R/computePairwiseMI.R:94-97."""
    nsnp = len(POS)
    snp_subset = int(min(nsnp, np.round(nsnp * 0.1)))  # R round(): half-to-even
    idx = RMersenne(seed).sample(nsnp, snp_subset)
    P = np.asarray(POS, dtype=np.float64)
    Ps = np.sort(P)
    total = 0
    for x in P[idx - 1]:
        # count of SNPs with circular distance > sr_dist  (0.5 g - |((x - P) mod g) - 0.5 g| > sr_dist)
        total += int(np.count_nonzero((0.5 * g - np.abs(np.mod(x - Ps, g) - 0.5 * g)) > sr_dist))
    return total / snp_subset * nsnp / 2
