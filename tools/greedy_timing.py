#!/usr/bin/env python
"""Time the greedy partition (kp_greedy) on synthetic data: python tools/greedy_timing.py [gen_pat ...]
Patterns beyond the DP's reach run on a lattice-free plan with k-mer counts drawn directly (no k-mer strings)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kmerpapa_b200.algorithms import greedy_penalty_plus_pseudo as gr
from kmerpapa_b200.engine import PartitionPlan

for gen_pat in (sys.argv[1:] or ["NNNANNN", "NNNNANNNN", "NNNNNANNNNN", "NNNNNNANNNNNN"]):
    plan = PartitionPlan(gen_pat, 0, lite=True)
    n = plan.nkmer
    rng = np.random.default_rng(9003)
    U = 1 + rng.negative_binomial(2, 2 / (2 + 33000.0 * 65536 / n), size=n)
    idx = np.arange(n)
    lograte = np.log(1e-3) + sum(rng.normal(0, s, 4)[(idx >> (2 * i)) & 3] for i, s in enumerate([0.5, 0.25, 0.12, 0.06, 0.03] + [0.0] * 16) if 4 ** i < n)
    M = rng.binomial(U, np.minimum(0.5, np.exp(lograte)))
    kM, kU = plan.upload_kmer_tables(M, U, name="gt")
    mu = M.sum() / (M.sum() + U.sum())
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        pats, loss, _, score = gr._greedy(plan, kM, kU, 1.0, (1 - mu) / mu, 6.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"{gen_pat}: {n} k-mers, {plan.npat:.3g} patterns: greedy partition of {len(pats)} patterns, score {score:.3f}, {dt * 1e3:.1f} ms", flush=True)
