#!/usr/bin/env python
"""How the DP time depends on the data regime (the work is data-oblivious except for the lazily computed exact scores):
python tools/regime_timing.py [gen_pat] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import get_plan

gen_pat = sys.argv[1] if len(sys.argv) > 1 else "NNNNANNNN"
plan = get_plan(gen_pat, 0)
rows = []
for name, mean_bg, rate, penalty in (("benchmark (bg 33000, rate 1e-3)", 33000.0, 1e-3, 6.0), ("sparse (bg 300, rate 1e-3)", 300.0, 1e-3, 6.0),
                                     ("very sparse (bg 30, rate 1e-2)", 30.0, 1e-2, 6.0), ("dense (bg 3e6, rate 1e-2)", 3e6, 1e-2, 6.0),
                                     ("huge penalty (everything merges)", 33000.0, 1e-3, 1e6), ("zero penalty", 33000.0, 1e-3, 0.0)):
    kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003, mean_bg=mean_bg, base_rate=rate)
    kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
    eM, eU = plan.expand(kM, kU)
    mc = int(pos.sum() + neg.sum())
    mu = int(pos.sum()) / mc
    beta = 1.0 * (1 - mu) / mu
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        best, kept = plan.dp_single(eM, eU, mc, 1.0, beta, penalty)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    part = plan.backtrack(best, kept)
    rows.append({"regime": name, "mean_background_per_kmer": mean_bg, "base_rate": rate, "penalty": penalty, "dp_ms": round(min(ts), 3),
                 "pattern_scores_per_s": plan.npat / min(ts) * 1e3, "roofline_frac_41B": 41.0 * plan.npat / (min(ts) / 1e3) / 1e9 / 6550.4,
                 "partition_patterns": int(len(part)), "wide_counts": bool(mc > 0xFFFFFFFF)})
    print(f"{name:36s} {min(ts):8.3f} ms  {plan.npat / min(ts) / 1e6:7.1f} Gpat/s  partition {len(part)} patterns", flush=True)
if len(sys.argv) > 2:
    json.dump({"gen_pat": gen_pat, "kernel": plan.dp_kernel_name(), "what": "K3+K4 time of one DP per data regime, best of 3, CUDA events",
               "rows": rows}, open(sys.argv[2], "w"), indent=1)
