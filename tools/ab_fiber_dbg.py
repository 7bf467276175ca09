#!/usr/bin/env python
"""Timing experiments on the fiber kernel: KP_FIBER_DBG = 0 (full), 1 (no stream phase), 2 (no level phase), 3 (neither).
Results are wrong for dbg != 0; only the times mean something.  python tools/ab_fiber_dbg.py [gen_pat] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import PartitionPlan

gen_pat = sys.argv[1] if len(sys.argv) > 1 else "NNNNANNNN"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003)
os.environ["KP_DP_KERNEL"] = "fiber"
plans = {}
for dbg in (0, 1, 2, 3):
    os.environ["KP_FIBER_DBG"] = str(dbg)
    plans[dbg] = PartitionPlan(gen_pat, 0)
plan = plans[0]
kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum())
mu = int(pos.sum()) / mc
best = plan._buffer("best", int(plan.info.table_elems), torch.float32)
kept = plan._buffer("kept", int(plan.info.kept_elems), torch.int16)
for dbg in (0, 1, 2, 3, 0):
    p = plans[dbg]
    p._buf["best"], p._buf["kept"] = best, kept
    ts = []
    for rep in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.dp_single(eM, eU, mc, 1.0, (1 - mu) / mu, 6.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"dbg={dbg}: min {min(ts[1:]):.3f} median {np.median(ts[1:]):.3f} ms", flush=True)
