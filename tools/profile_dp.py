#!/usr/bin/env python
"""One 9-mer single DP (and optionally one CV job) for ncu: `python tools/profile_dp.py [single|cv] [gen_pat]`.
`reps` DPs (the first one warms up); the kernel of interest is kp_dp_rows_kernel (16 launches per DP, one per wave)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import get_plan

kind = sys.argv[1] if len(sys.argv) > 1 else "single"
gen_pat = sys.argv[2] if len(sys.argv) > 2 else "NNNNANNNN"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003)
codes = synthetic.codes_of(kmers)
plan = get_plan(gen_pat, 0)
kM, kU = plan.pack_counts(codes, pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum())
mu = int(pos.sum()) / mc
beta = 1.0 * (1 - mu) / mu
for rep in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if kind == "single":
        best, kept = plan.dp_single(eM, eU, mc, 1.0, beta, 6.0)
    else:
        kMf, kUf = plan.upload_kmer_tables(pos // 5, neg // 5, name="pf")
        fM, fU = plan.expand(kMf, kUf, name="pfe")
        print(plan.cv_job(eM, eU, fM, fU, mc, 1.0, beta, 6.0))
    e1.record()
    torch.cuda.synchronize()
    print(f"{kind} {gen_pat} rep {rep}: {e0.elapsed_time(e1):.3f} ms, {plan.npat / e0.elapsed_time(e1) / 1e6:.2f} Gpat/s", flush=True)
