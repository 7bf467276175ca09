#!/usr/bin/env python
"""The command line on a full N^9 general pattern (38.4 G patterns: the score table, 169 GB, does not fit one GPU),
under torchrun on >= 2 GPUs: torchrun --nproc-per-node 2 tools/cli_n9_demo.py"""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from kmerpapa_b200 import cli

rank = int(os.environ.get("RANK", "0"))
gen_pat = "NNNNNNNNN"
n = 4 ** 9
rng = np.random.default_rng(9009)
bg = 1 + rng.negative_binomial(2, 2 / (2 + 8000.0), size=n)
idx = np.arange(n)
lograte = np.log(1e-3) + sum(rng.normal(0, s, 4)[(idx >> (2 * i)) & 3] for i, s in enumerate([0.03, 0.06, 0.12, 0.25, 0.5, 0.5, 0.25, 0.12, 0.06]))
pos = rng.binomial(bg, np.minimum(0.5, np.exp(lograte)))
letters = "ACGT"
d = f"/tmp/n9_{rank}"
os.makedirs(d, exist_ok=True)
with open(f"{d}/pos.txt", "w") as fp, open(f"{d}/bg.txt", "w") as fb:
    for i in range(n):
        k = "".join(letters[(i >> (2 * j)) & 3] for j in range(9))
        fp.write(f"{k} {pos[i]}\n")
        fb.write(f"{k} {bg[i]}\n")
t = time.perf_counter()
err = io.StringIO()
with contextlib.redirect_stderr(err):
    rc = cli.main(["-p", f"{d}/pos.txt", "-b", f"{d}/bg.txt", "-c", "8", "-a", "1", "-o", f"{d}/out.txt"])
dt = time.perf_counter() - t
if rank == 0:
    print(err.getvalue().strip())
    print(f"rc={rc}, {dt:.1f} s end to end (parse, plan, pack, expand, sharded DP, backtrack, counts, output)")
    print(open(f"{d}/out.txt").read()[:400])
import torch.distributed as dist

if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
