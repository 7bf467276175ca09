import os, sys
sys.path.insert(0, os.getcwd())
import torch
from kmerpapa_b200 import synthetic
from kmerpapa_b200.engine import get_plan
gen_pat = "NNNNANNNN"
kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9003, mean_bg=300.0, base_rate=1e-3)
plan = get_plan(gen_pat, 0)
kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum()); mu = int(pos.sum()) / mc
for rep in range(2):
    plan.dp_single(eM, eU, mc, 1.0, (1 - mu) / mu, 6.0)
torch.cuda.synchronize()
