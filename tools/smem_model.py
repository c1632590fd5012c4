"""Offline model of the shared-memory traffic of offt_b200/csrc/fft_kernels.cuh.

Mirrors the kernel's index algebra exactly and counts shared-memory wavefronts per
request (1.0 = conflict-free) so that the PAD constants of each FftCfg can be chosen
without a GPU.  `python tools/smem_model.py` prints the best pads for every config.
"""
import itertools
import sys


def brev(v, radix):
    r, m = 0, radix >> 1
    while m:
        r = (r << 1) | (v & 1); v >>= 1; m >>= 1
    return r


class Cfg:
    def __init__(self, N, E, radices, pads):
        self.N, self.E, self.R, self.pads = N, E, [r for r in radices if r > 1] or [radices[0]], list(pads) + [0, 0, 0]
        self.T = N // E
        self.NS = len(self.R)

    def P(self, s):
        p = 1
        for j in range(s):
            p *= self.R[j]
        return p

    def M(self, s):
        return self.N // (self.P(s) * self.R[s])

    def pitch(self, s):
        return self.N // self.R[s + 1] + self.pads[s]

    def colsize(self):
        m = self.N + 1
        for s in range(self.NS - 1):
            m = max(m, self.R[s + 1] * self.pitch(s))
        return m


def requests(cfg, C, cfast):
    """yield (kind, [address per thread of the CTA]) for every shared-memory instruction"""
    nthreads = cfg.T * C
    def tc(tid):
        return (tid // C, tid % C) if cfast else (tid % cfg.T, tid // cfg.T)
    def sidx(A, c):
        return A * C + c if cfast else c * cfg.colsize() + A
    for s in range(cfg.NS):
        R, NU = cfg.R[s], cfg.E // cfg.R[s]
        if s > 0:
            for u in range(NU):
                for i in range(R):
                    yield ("ld", [sidx(i * cfg.pitch(s - 1) + tc(tid)[0] + cfg.T * u, tc(tid)[1]) for tid in range(nthreads)])
        if s < cfg.NS - 1:
            P, M = cfg.P(s), cfg.M(s)
            Rn = cfg.R[s + 1]; Mn = M // Rn
            for u in range(NU):
                for pos in range(R):
                    k = brev(pos, R)
                    addrs = []
                    for tid in range(nthreads):
                        t, c = tc(tid)
                        beta = t + cfg.T * u
                        np_, K = beta % M, beta // M
                        npp, inext = np_ % Mn, np_ // Mn
                        addrs.append(sidx(inext * cfg.pitch(s) + npp + Mn * K + Mn * P * k, c))
                    yield ("st", addrs)


def wavefronts(addrs, G):
    """average wavefronts per G-lane group (G = 128 B / element size)"""
    tot = n = 0
    for g0 in range(0, len(addrs), G):
        grp = addrs[g0:g0 + G]
        banks = {}
        for a in set(grp):
            banks[a % G] = banks.get(a % G, 0) + 1
        tot += max(banks.values()); n += 1
    return tot / n


def score(cfg, C, cfast, G):
    w = [wavefronts(a, G) for _, a in requests(cfg, C, cfast)]
    return sum(w) / len(w) if w else 1.0


CONFIGS = {  # N: (E, radices)
    2: (2, (2,)), 4: (4, (4,)), 8: (8, (8,)), 16: (16, (16,)), 32: (8, (8, 4)), 64: (8, (8, 8)),
    128: (16, (16, 8)), 256: (16, (16, 16)), 512: (8, (8, 8, 8)), 1024: (16, (16, 16, 4)),
    2048: (16, (16, 16, 8)), 4096: (16, (16, 16, 16)), 8192: (16, (16, 16, 16, 2)),
}

if __name__ == "__main__":
    for N, (E, rad) in CONFIGS.items():
        ns = len(rad)
        if ns == 1:
            continue
        for G, name in ((8, "c128"), (16, "c64")):
            best = None
            for pads in itertools.product(range(0, G + 1), repeat=ns - 1):
                cfg = Cfg(N, E, rad, pads)
                C = max(1, min(256 // cfg.T, 8))
                sc = score(cfg, C, False, G)
                # also look at the strided launch with few columns
                sc2 = score(cfg, min(4, G), True, G)
                key = (round(sc, 3), round(sc2, 3), sum(pads))
                if best is None or key < best[0]:
                    best = (key, pads)
            print(f"N={N:5d} E={E:2d} radices={rad} {name}: pads={best[1]} n-fast wavefronts/request={best[0][0]} c-fast(C=4)={best[0][1]}")
