import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import get_plan
gen_pat = "NNNNANNNN"
kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003)
plan = get_plan(gen_pat, 0)
kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum()); mu = int(pos.sum()) / mc; beta = (1 - mu) / mu
res = {0: [], 1: []}
for rep in range(int(os.environ.get("AB_REPS", "24"))):
    mode = rep & 1
    if mode: os.environ.pop("KP_ONE_LAUNCH", None)
    else: os.environ["KP_ONE_LAUNCH"] = "1"
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.dp_single(eM, eU, mc, 1.0, beta, 6.0); e1.record(); torch.cuda.synchronize()
    if rep >= 4: res[mode].append(e0.elapsed_time(e1))
for m in (0, 1):
    a = np.array(res[m]); print("single-launch" if m == 0 else "wave launches", "min %.3f median %.3f max %.3f" % (a.min(), np.median(a), a.max()))
