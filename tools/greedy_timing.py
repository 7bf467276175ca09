#!/usr/bin/env python
"""Time the greedy partition (kp_greedy) on synthetic data: python tools/greedy_timing.py [gen_pat]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kmerpapa_b200 import synthetic
from kmerpapa_b200.algorithms import greedy_penalty_plus_pseudo as gr
from kmerpapa_b200.engine import get_plan

for gen_pat in (sys.argv[1:] or ["NNNANNN", "NNNNANNNN"]):
    kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003)
    plan = get_plan(gen_pat, 0)
    kM, kU = plan.upload_kmer_tables(pos, neg, name="gt")
    mu = pos.sum() / (pos.sum() + neg.sum())
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        pats, loss, _, score = gr._greedy(plan, kM, kU, 1.0, (1 - mu) / mu, 6.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"{gen_pat}: greedy partition {len(pats)} patterns, score {score:.3f}, {dt * 1e3:.2f} ms", flush=True)
