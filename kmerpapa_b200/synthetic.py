"""Synthetic k-mer count tables for the benchmark workloads (BASELINE configs 3-5).

Negative-binomial background counts and a position-dependent log-linear mutation rate.  The DP's work
is data-oblivious (every split of every pattern is evaluated), so the distribution only shapes the
resulting partition, not the throughput.
"""
import numpy as np

from . import iupac


def position_sigmas(k):
    """Effect size per position: 0 at the centre, 0.5 next to it, then 0.25, 0.12, 0.06, 0.03 outward."""
    steps = [0.5, 0.25, 0.12, 0.06, 0.03]
    c = k // 2
    out = []
    for i in range(k):
        d = abs(i - c)
        out.append(0.0 if d == 0 else steps[min(d, len(steps)) - 1])
    return out


def negbin_counts(gen_pat, seed, mean_bg=33000.0, base_rate=1e-3):
    """Returns (kmers, pos, neg) for every k-mer of gen_pat; k-mers in k-mer index order, counts drawn
    over the k-mers in sorted order (so the table does not depend on the enumeration order)."""
    kmers = iupac.matches(gen_pat)
    order = np.argsort(np.array(kmers))
    rng = np.random.default_rng(seed)
    n, k = len(kmers), len(gen_pat)
    bg_sorted = 1 + rng.negative_binomial(2, 2.0 / (2.0 + mean_bg), size=n)
    effects = [rng.normal(0.0, s, size=4) if s > 0 else np.zeros(4) for s in position_sigmas(k)]
    base_index = {"A": 0, "C": 1, "G": 2, "T": 3}
    lograte = np.full(n, np.log(base_rate))
    sorted_kmers = [kmers[i] for i in order]
    for i in range(k):
        col = np.array([base_index[km[i]] for km in sorted_kmers])
        lograte += effects[i][col]
    rate = np.minimum(0.5, np.exp(lograte))
    pos_sorted = rng.binomial(bg_sorted, rate)
    pos = np.empty(n, dtype=np.int64)
    bg = np.empty(n, dtype=np.int64)
    pos[order] = pos_sorted
    bg[order] = bg_sorted
    return kmers, pos, bg - pos


def codes_of(kmers):
    return np.array([iupac.kmer_code(km) for km in kmers], dtype=np.uint64)
