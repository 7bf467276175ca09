#!/usr/bin/env python
"""Interleaved A/B of one environment knob of the DP (knobs are read once, at plan creation: one plan per setting):
python tools/ab_env.py NAME VALUE_A VALUE_B [gen_pat] [reps]   ('-' = unset)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import PartitionPlan

name, va, vb = sys.argv[1:4]
gen_pat = sys.argv[4] if len(sys.argv) > 4 else "NNNNANNNN"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 30
kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003)
plans = []
for v in (va, vb):
    if v == "-":
        os.environ.pop(name, None)
    else:
        os.environ[name] = v
    plans.append(PartitionPlan(gen_pat, 0))
plan = plans[0]
kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum())
mu = int(pos.sum()) / mc
res = {0: [], 1: []}
for rep in range(reps + 4):
    mode = rep & 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plans[mode].dp_single(eM, eU, mc, 1.0, (1 - mu) / mu, 6.0)
    e1.record()
    torch.cuda.synchronize()
    if rep >= 4:
        res[mode].append(e0.elapsed_time(e1))
for m, v in ((0, va), (1, vb)):
    a = np.array(res[m])
    print(f"{name}={v}: min {a.min():.3f} median {np.median(a):.3f} max {a.max():.3f} ms")
