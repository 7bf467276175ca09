#!/usr/bin/env python
"""Soak test of the DP: the same DP many times, the whole score table and the kept flags must come out bit-identical
every time.  Default: the opt-in single-launch mode (tiles wait on completion flags of other tiles); KP_ONE_LAUNCH=0
in the environment soaks the default one-launch-per-wave mode.  python tools/soak_dp.py [reps] [gen_pat]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("KP_ONE_LAUNCH", "1")   # the mode under test (opt-in)
import torch

from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import get_plan

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gen_pat = sys.argv[2] if len(sys.argv) > 2 else "NNNNANNNN"
kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003)
plan = get_plan(gen_pat, 0)
kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum())
mu = int(pos.sum()) / mc


def run(alpha, penalty):
    best, kept = plan.dp_single(eM, eU, mc, alpha, alpha * (1 - mu) / mu, penalty)
    a = best.view(torch.int32)
    # two independent checksums of the table bits, and one of the flags
    return (int(a.sum(dtype=torch.int64).item()), int((a.to(torch.int64) * 2654435761 % 4294967291).sum().item()),
            int(kept.to(torch.int64).sum().item()), plan.top_score(best).tobytes())


settings = [(1.0, 6.0), (0.5, 3.0), (10.0, 5.0)]
ref = {}
bad = 0
for i in range(reps):
    s = settings[i % len(settings)]
    r = run(*s)
    if s not in ref:
        ref[s] = r
    elif r != ref[s]:
        bad += 1
        print("MISMATCH at iteration", i, s, r, ref[s], flush=True)
print(f"soak {gen_pat}: {reps} DPs, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
