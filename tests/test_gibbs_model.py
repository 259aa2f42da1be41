"""The argument behind the Gibbs kernels (rambl_b200/csrc/dpm.cu), checked on the CPU with a
small model: a round of B draws is picked SPECULATIVELY with the masses at the start of the round, then every draw
checks its pick against the picks currently believed for the earlier draws and re-derives it when the check fails,
all draws at once, pass after pass.  Claim: the fixed point is the sequential chain (draw t sees the masses left by
draws < t: NonparametricClustering.cpp:171-196), it is reached after at most B+1 passes, and draw j is final after
pass 1 and j check passes.  The model uses the kernels' definition of a cumulative weight,

    cum_j(s) = base_j(s) + corr_j(s),   base_j(s) = sum_{s'<=s} m[s'] w_j[s']   (m = masses at the start of the round)
                                        corr_j(s) = sum_{s'<=s} k_j[s'] w_j[s']  (k_j = earlier picks of s' in the round)

with both sums taken in strain order, so "sequential" and "speculative" are compared bit for bit.  This is a test of
the algorithm, not of the CUDA code: the -m gpu tests compare the kernels with the oracle and with each other."""
import numpy as np
import pytest


def _pick(base, corr, thr):
    """std::discrete_distribution: first strain of 0..S-2 whose cumulative weight reaches thr, else S-1."""
    S = len(base)
    for s in range(S - 1):
        if not (base[s] + corr[s] < thr):
            return s
    return S - 1


def _prefix(values):
    out, acc = [], 0.0
    for v in values:
        acc = acc + v
        out.append(acc)
    return out


def _corr(w_j, counts):
    return _prefix([counts[s] * w_j[s] for s in range(len(w_j))])


def sequential_chain(w, m0, u, B):
    """Draw by draw; the masses move at the end of each round of B draws, the picks inside a round enter as counts."""
    T, S = w.shape
    picks, m = [], list(m0)
    for r0 in range(0, T, B):
        counts = [0] * S
        for t in range(r0, min(T, r0 + B)):
            base, corr = _prefix([m[s] * w[t][s] for s in range(S)]), _corr(w[t], counts)
            c = _pick(base, corr, u[t] * (base[-1] + corr[-1]))
            picks.append(c)
            counts[c] += 1
        m = [m[s] + counts[s] for s in range(S)]
    return picks


def speculative_chain(w, m0, u, B):
    """Rounds of B draws: pass 1 without corrections, then check-and-re-derive passes until nothing moves."""
    T, S = w.shape
    picks, m, most_passes, late_moves = [], list(m0), 0, 0
    for r0 in range(0, T, B):
        n = min(T, r0 + B) - r0
        base = [_prefix([m[s] * w[r0 + j][s] for s in range(S)]) for j in range(n)]
        zero = [0.0] * S
        c = [_pick(base[j], zero, u[r0 + j] * base[j][-1]) for j in range(n)]
        passes = 1
        while True:
            published = list(c)  # every draw sees the picks of the previous pass
            moved = False
            for j in range(n):
                counts = [0] * S
                for i in range(j):
                    counts[published[i]] += 1
                corr = _corr(w[r0 + j], counts)
                thr = u[r0 + j] * (base[j][-1] + corr[-1])
                cj = published[j]
                lo_ok = cj == 0 or base[j][cj - 1] + corr[cj - 1] < thr
                hi_ok = cj == S - 1 or not (base[j][cj] + corr[cj] < thr)
                if not (lo_ok and hi_ok):
                    new = _pick(base[j], corr, thr)
                    if new != cj:
                        moved = True
                        if passes > j:
                            late_moves += 1  # draw j is final after pass 1 and j check passes
                        c[j] = new
            passes += 1
            if not moved:
                break
        most_passes = max(most_passes, passes)
        assert passes <= n + 2
        picks.extend(c)
        counts = [0] * S
        for x in c:
            counts[x] += 1
        m = [m[s] + counts[s] for s in range(S)]
    return picks, most_passes, late_moves


@pytest.mark.parametrize("seed,S,B,T,mass_scale", [
    (0, 2, 32, 200, 1.0), (1, 5, 32, 333, 0.01), (2, 17, 64, 400, 0.001), (3, 48, 128, 700, 0.1),
    (4, 33, 256, 900, 0.01), (5, 65, 96, 500, 1e-4), (6, 100, 160, 640, 0.05), (7, 3, 256, 1000, 1e-6),
])
def test_speculative_rounds_settle_on_the_sequential_chain(seed, S, B, T, mass_scale):
    rng = np.random.default_rng(seed)
    # weights like exp(log-likelihood): a few strains explain a read, many are (almost) impossible, some exactly 0
    w = np.exp(-rng.integers(0, 12, size=(T, S)) * rng.random((T, S)) * 3.0)
    w[rng.random((T, S)) < 0.1] = 0.0
    w[np.arange(T), rng.integers(0, S, size=T)] = 1.0  # every read has a strain that explains it
    m0 = (rng.random(S) + 0.05) * mass_scale          # tiny masses: the first picks of a round move the later ones
    u = rng.random(T)
    want = sequential_chain(w, m0, u, B)
    got, most_passes, late_moves = speculative_chain(w, m0, u, B)
    assert got == want
    assert late_moves == 0
    if mass_scale <= 0.01:
        assert most_passes > 2  # the case is hard enough to need re-derivations
