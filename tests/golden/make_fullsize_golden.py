#!/usr/bin/env python
"""Full-size goldens from the CPU oracle (oracle/kp_oracle.c) -> tests/golden/fullsize.json.

The oracle is pinned bit for bit to the unmodified reference on every fixture under tests/golden/ (tests/test_oracle.py);
the reference itself (numba, one core) needs 25-70 min and up to 51 GB per full-size DP and cannot run config 4 at all
(SURVEY 8c/8d), so at the sizes of BASELINE configs 3-5 the goldens come from the oracle.  This script needs a host with a
lot of memory (config 3: 54 GB for one DP; config 4: 102 GB per job), so it is run once on the GPU box's host cores
(`gpurun -- python tests/golden/make_fullsize_golden.py --out gpurun_out/fullsize.json cfg3 cfg5 cfg4`) and its output is
committed.  It uses no GPU.

Per single DP (cfg3, cfg5): loss (float32 bits), partition (count, SHA-256 of the dense pattern numbers as little-endian
uint64 in emission order), number of kept-whole patterns, sum and sum of squares (mod 2^64) of the float32 score bits of
the whole table.  cfg4: train and held-out float32 loss of the general pattern for every (fold, alpha, penalty) job that
was run, and the selection when all 45 were.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

ALPHA, PENALTY = 1.0, 6.0
SINGLE = {"cfg3": ("NNNNANNNN", 9003), "cfg5": ("RYNNNANNNRY", 9005)}
CV_ALPHAS, CV_PENALTIES, CV_FOLDS, CV_SEED = [0.5, 1.0, 10.0], [3.0, 5.0, 6.0], 5, 1


def checksums(score):
    bits = score.view(np.uint32)
    s1 = s2 = 0
    step = 1 << 26
    for lo in range(0, bits.size, step):
        b = bits[lo:lo + step].astype(np.uint64)
        s1 = (s1 + int(b.sum(dtype=np.uint64))) & ((1 << 64) - 1)
        s2 = (s2 + int((b * b).sum(dtype=np.uint64))) & ((1 << 64) - 1)
    return s1, s2


def single(name, O, threads):
    from kmerpapa_b200 import synthetic

    gen_pat, seed = SINGLE[name]
    kmers, pos, neg = synthetic.negbin_counts(gen_pat, seed)
    mc = int(pos.sum() + neg.sum())
    mu = int(pos.sum()) / mc
    beta = (ALPHA * (1.0 - mu)) / mu
    t0 = time.time()
    res = O.single_dp(gen_pat, pos, neg, ALPHA, beta, PENALTY, nthreads=threads)
    t1 = time.time()
    pat = O.backtrack(gen_pat, res["split"])
    s1, s2 = checksums(res["score"])
    kept = 0
    step = 1 << 28
    for lo in range(0, res["split"].size, step):
        kept += int((res["split"][lo:lo + step] == 0xFF).sum())
    loss = res["score"][-1]
    return {"gen_pat": gen_pat, "seed": seed, "alpha": ALPHA, "penalty": PENALTY, "beta": beta, "npat": int(res["score"].size),
            "loss": float(loss), "loss_bits": f"{int(loss.view(np.uint32)):08x}", "partition_patterns": int(len(pat)),
            "partition_sha256": hashlib.sha256(np.ascontiguousarray(pat, dtype="<u8").tobytes()).hexdigest(),
            "kept_whole": kept, "score_bits_sum": str(s1), "score_bits_sumsq": str(s2),
            "oracle_seconds": round(t1 - t0, 2), "threads": threads}


def cv(O, threads, jobs, previous):
    from kmerpapa_b200 import CV_tools, synthetic
    from kmerpapa_b200.score_utils import get_betas

    gen_pat = "NNNNANNNN"
    kmers, pos, neg = synthetic.negbin_counts(gen_pat, 9004)
    Mf, Uf = CV_tools.sample_fold_counts(kmers, pos, neg, CV_FOLDS, np.random.RandomState(CV_SEED))
    Mtot, Utot = Mf.sum(axis=1), Uf.sum(axis=1)
    M_train = Mf.sum() - Mf.sum(axis=0)
    U_train = Uf.sum() - Uf.sum(axis=0)
    npat, _, _ = O.plan_info(gen_pat)
    shape = (1, CV_FOLDS, len(CV_ALPHAS), len(CV_PENALTIES), 2)
    bits = np.zeros(shape, dtype=np.uint32)
    done = set()
    if previous:
        bits = np.array([int(x, 16) for x in previous["job_bits"]], dtype=np.uint32).reshape(shape)
        done = set(map(tuple, previous["jobs_run"]))
    secs = []
    for f, a_i, p_i in jobs:
        if (f, a_i, p_i) in done:
            continue
        alpha, penalty = CV_ALPHAS[a_i], CV_PENALTIES[p_i]
        beta = get_betas(alpha, M_train, U_train)[f]
        t0 = time.time()
        train, test = O.cv_job(gen_pat, Mtot, Utot, Mf[:, f], Uf[:, f], alpha, beta, penalty, nthreads=threads)
        secs.append(time.time() - t0)
        bits[0, f, a_i, p_i, 0] = train[npat - 1].view(np.uint32)
        bits[0, f, a_i, p_i, 1] = test[npat - 1].view(np.uint32)
        del train, test
        done.add((f, a_i, p_i))
        print(f"cfg4 job fold={f} alpha={alpha} penalty={penalty}: {secs[-1]:.1f} s", file=sys.stderr, flush=True)
    out = {"gen_pat": gen_pat, "seed": 9004, "alphas": CV_ALPHAS, "penalties": CV_PENALTIES, "nfolds": CV_FOLDS, "cv_seed": CV_SEED,
           "job_bits": [f"{int(x):08x}" for x in bits.reshape(-1)], "jobs_run": sorted(map(list, done)),
           "oracle_seconds_per_job": round(float(np.mean(secs)), 2) if secs else (previous or {}).get("oracle_seconds_per_job"),
           "threads": threads}
    if len(done) == CV_FOLDS * len(CV_ALPHAS) * len(CV_PENALTIES):
        from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cvmod

        a, c, t = cvmod.select_best(CV_ALPHAS, CV_PENALTIES, bits.view(np.float32), 1, CV_FOLDS, len(gen_pat))
        out["selected"] = [a, c, float(t)]
        out["selected_bits"] = f"{int(np.float32(t).view(np.uint32)):08x}"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="+", choices=["cfg3", "cfg5", "cfg4"])
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "fullsize.json"))
    ap.add_argument("--merge", default=os.path.join(ROOT, "tests", "golden", "fullsize.json"), help="start from this file")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--cv-jobs", type=int, default=45, help="how many of the 45 jobs to run (spread over folds and grid points)")
    args = ap.parse_args()
    from oracle import kp_oracle as O

    O.build()
    out = json.load(open(args.merge)) if os.path.exists(args.merge) else {}
    for name in args.what:
        if name == "cfg4":
            alljobs = [(f, a, p) for f in range(CV_FOLDS) for a in range(3) for p in range(3)]
            # a spread first (every fold, alpha and penalty appears early), then the rest
            order = sorted(alljobs, key=lambda j: ((j[0] + j[1] + j[2]) % 5 != 0, (j[0] * 3 + j[1] + 2 * j[2]) % 7, j))
            out[name] = cv(O, args.threads, order[: args.cv_jobs], out.get(name))
        else:
            out[name] = single(name, O, args.threads)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
        print(name, json.dumps({k: v for k, v in out[name].items() if k != "job_bits"}), file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
