#!/usr/bin/env python
"""Times the UNMODIFIED reference (kmerPaPa v0.2.4, numba) on one single DP, in a fresh process.

BASELINE / TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and --impl reference).  The reference package is imported
from baseline/_ref (a plain copy of the reference's pure-Python package made by __graft_entry__.build() in the build
container; git-ignored, shipped to the GPU box), never from /root/reference.  One process per call: numba freezes
alpha / beta / penalty as compile-time globals at the first call, so a process can time exactly one DP.

  python oracle/ref_numba_timing.py GEN_PAT SEED ALPHA PENALTY

Times `kmerpapa.algorithms.bottum_up_array_w_numba.pattern_partition_bottom_up` (bottum_up_array_w_numba.py:67-124) on the
synthetic counts of kmerpapa_b200.synthetic.negbin_counts(GEN_PAT, SEED); the time includes numba's JIT (bench.py
measures that share with a second call of this script on a tiny general pattern).  Prints one JSON line.
"""
import argparse
import hashlib
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    gen_pat, seed, alpha, penalty = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4])
    if not os.path.isdir(os.path.join(REF, "kmerpapa")):
        print(json.dumps({"unavailable": "baseline/_ref/kmerpapa is missing (run __graft_entry__.build() where /root/reference exists)"}))
        return
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import numpy as np

    from kmerpapa_b200 import synthetic

    kmers, pos, neg = synthetic.negbin_counts(gen_pat, seed)
    n_mut, n_unmut = int(pos.sum()), int(neg.sum())
    mu = n_mut / (n_mut + n_unmut)
    beta = (alpha * (1.0 - mu)) / mu
    contextD = {k: (int(p), int(u)) for k, p, u in zip(kmers, pos, neg)}
    t_imp = time.perf_counter()
    import kmerpapa.algorithms.bottum_up_array_w_numba as ref   # the reference's own module, unmodified
    import kmerpapa.pattern_utils as pu

    args = argparse.Namespace(verbosity=0)
    t0 = time.perf_counter()
    score, M, U, names = ref.pattern_partition_bottom_up(gen_pat, contextD, alpha, beta, penalty, args, n_mut, n_unmut)
    dt = time.perf_counter() - t0
    import numba

    print(json.dumps({
        "gen_pat": gen_pat, "seed": seed, "alpha": alpha, "penalty": penalty, "npat": int(pu.pattern_max(gen_pat)),
        "seconds": dt, "import_seconds": t0 - t_imp, "loss": float(score),
        "loss_bits": f"{int(np.float32(score).view(np.uint32)):08x}", "patterns": len(names),
        "partition_sha256": hashlib.sha256("\n".join(names).encode()).hexdigest(),
        "numba": numba.__version__, "numpy": np.__version__, "threads": 1}))


if __name__ == "__main__":
    main()
