#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and numba); the fixtures it writes are
committed, so nothing in tests/, smoke() or bench.py needs the reference at run time.

Each reference call runs in a fresh subprocess: the single-DP module freezes alpha/beta/penalty
as numba globals at first compile, so a second in-process call is unreliable.

Fixtures
  single_<name>.npz   full tables of bottum_up_array_w_numba.pattern_partition_bottom_up
                      (score float32, M, U, backtrack pointer, partition names) on small general
                      patterns; the arrays are captured by wrapping numpy.full/numpy.empty inside
                      the worker, the reference code itself is untouched.
  cv_<name>.npz       bottum_up_array_penalty_plus_pseudo_CV.pattern_partition_bottom_up: the
                      float32 train table of every grid point, the held-out fold counts, CVfile
                      rows and the selected (alpha, penalty).  numpy.empty is mapped to numpy.zeros
                      in the worker because the reference sums never-written rows of np.empty
                      tables; zero pages are what large runs see.
  cli_<name>.json     stdout / CVfile / stderr of the reference command line on test_data
                      (BASELINE configs 1 and 2) and on small synthetic inputs.

Usage:  python tests/golden/make_golden.py [all|small|cli5|cli7|...]
"""
import json
import os
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
REF_DATA = "/root/reference/test_data"

CODE = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "GC", "W": "AT", "K": "GT",
        "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}


def kmers_of(gen_pat):
    """k-mers of gen_pat, first position fastest (same order as the reference's matches())."""
    out = [""]
    for ch in reversed(gen_pat):
        out = [b + s for s in out for b in CODE[ch]]
    return out


# ---------------------------------------------------------------------------------------------
# worker side (fresh process per call)
# ---------------------------------------------------------------------------------------------
def _stub_skopt():
    sk = types.ModuleType("skopt")
    sk.gp_minimize = None
    sp = types.ModuleType("skopt.space")
    sp.Real = sp.Integer = object
    su = types.ModuleType("skopt.utils")
    su.use_named_args = lambda *a, **k: (lambda f: f)
    sys.modules.update({"skopt": sk, "skopt.space": sp, "skopt.utils": su})


def worker_single(spec_path, out_path):
    import numpy as np

    spec = json.load(open(spec_path))
    captured = []
    real_full, real_empty = np.full, np.empty

    def cap_full(*a, **k):
        arr = real_full(*a, **k)
        captured.append(("full", arr))
        return arr

    def cap_empty(*a, **k):
        arr = real_empty(*a, **k)
        captured.append(("empty", arr))
        return arr

    sys.path.insert(0, REF_SRC)
    from kmerpapa.algorithms import bottum_up_array_w_numba as ref

    gen_pat = spec["gen_pat"]
    kmers = kmers_of(gen_pat)
    contextD = {km: (int(m), int(u)) for km, m, u in zip(kmers, spec["M"], spec["U"])}
    nmut, nunmut = sum(spec["M"]), sum(spec["U"])
    args = types.SimpleNamespace(verbosity=0)
    np.full, np.empty = cap_full, cap_empty
    try:
        score, M, U, names = ref.pattern_partition_bottom_up(
            gen_pat, contextD, spec["alpha"], spec["beta"], spec["penalty"], args, nmut, nunmut)
    finally:
        np.full, np.empty = real_full, real_empty
    npat = ref.pattern_max(gen_pat)
    # allocation order in the reference: score_mem (full), U_mem, M_mem, backtrack_mem (empty)
    tabs = [a for kind, a in captured if getattr(a, "shape", None) == (npat,)]
    score_mem = next(a for a in tabs if a.dtype == np.float32)
    ints = [a for a in tabs if a.dtype in (np.uint32, np.uint64)]
    U_mem, M_mem, bt = ints[0], ints[1], ints[2]
    assert bt.dtype == np.uint64
    np.savez_compressed(
        out_path, gen_pat=gen_pat, alpha=spec["alpha"], beta=spec["beta"], penalty=spec["penalty"],
        kmerM=np.array(spec["M"], dtype=np.uint64), kmerU=np.array(spec["U"], dtype=np.uint64),
        score=score_mem, M=M_mem.astype(np.uint64), U=U_mem.astype(np.uint64), bt=bt,
        top_score=np.float32(score), top_M=np.uint64(M), top_U=np.uint64(U), names=np.array(names))


def worker_cv(spec_path, out_path):
    import numpy as np

    spec = json.load(open(spec_path))
    real_full, real_empty = np.full, np.empty
    fulls, empties = [], []

    def cap_full(*a, **k):
        arr = real_full(*a, **k)
        fulls.append(arr)
        return arr

    def zeros_for_empty(shape, dtype=float, **k):
        arr = np.zeros(shape, dtype=dtype)
        empties.append(arr)
        return arr

    sys.path.insert(0, REF_SRC)
    from kmerpapa.algorithms import bottum_up_array_penalty_plus_pseudo_CV as ref

    gen_pat = spec["gen_pat"]
    kmers = kmers_of(gen_pat)
    contextD = {km: (int(m), int(u)) for km, m, u in zip(kmers, spec["M"], spec["U"])}
    nmut, nunmut = sum(spec["M"]), sum(spec["U"])
    cvfile = open(out_path + ".cvrows", "w")
    args = types.SimpleNamespace(verbosity=0, nfolds=spec["nfolds"], iterations=spec.get("iterations", 1),
                                 seed=spec["seed"], CVfile=cvfile)
    np.full, np.empty = cap_full, zeros_for_empty
    try:
        a, c, best = ref.pattern_partition_bottom_up(
            gen_pat, contextD, spec["alphas"], args, nmut, nunmut, spec["penalties"])
    finally:
        np.full, np.empty = real_full, real_empty
    cvfile.close()
    rows = open(out_path + ".cvrows").read()
    os.remove(out_path + ".cvrows")
    npat, nf = ref.pattern_max(gen_pat), spec["nfolds"]
    trains = [x for x in fulls if getattr(x, "shape", None) == (npat, nf) and x.dtype == np.float32]
    ints = [x for x in empties if x.shape == (npat, nf) and x.dtype in (np.uint32, np.uint64)]
    tests = [x for x in empties if x.shape == (npat, nf) and x.dtype == np.float32]
    U_mem, M_mem = ints[0], ints[1]  # allocation order in the reference: U_mem then M_mem
    np.savez_compressed(
        out_path, gen_pat=gen_pat, alphas=np.array(spec["alphas"]), penalties=np.array(spec["penalties"]),
        nfolds=nf, seed=spec["seed"], iterations=spec.get("iterations", 1),
        kmerM=np.array(spec["M"], dtype=np.uint64), kmerU=np.array(spec["U"], dtype=np.uint64),
        train_tables=np.stack(trains), last_test_table=tests[0],
        M_folds=M_mem.astype(np.uint64), U_folds=U_mem.astype(np.uint64),
        cv_rows=rows, best_alpha=float(a), best_penalty=float(c), best_test=np.float32(best))


def worker_cli(spec_path, out_path):
    import contextlib
    import io

    spec = json.load(open(spec_path))
    _stub_skopt()
    sys.path.insert(0, REF_SRC)
    from kmerpapa import cli

    argv = list(spec["argv"])
    out_file, cv_file = out_path + ".out", out_path + ".cv"
    argv += ["-o", out_file]
    if spec.get("cvfile", True):
        argv += ["--CVfile", cv_file]
    if spec.get("zero_empty"):
        # tiny CV tables (here the 3-mers of --test_smaller_k) come from recycled heap memory in the reference
        # and its fold totals then depend on that garbage; zero pages are what large runs see
        import numpy as np

        real_empty = np.empty
        np.empty = lambda shape, dtype=float, **k: np.zeros(shape, dtype=dtype)
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        rc = cli.main(argv)
    if spec.get("zero_empty"):
        np.empty = real_empty
    import gc

    gc.collect()   # cli.main leaves its output files to the garbage collector: make sure they are flushed
    res = {"argv": spec["argv"], "rc": rc, "stdout": open(out_file).read(),
           "cvfile": open(cv_file).read() if os.path.exists(cv_file) else None,
           "stderr": [l for l in err.getvalue().splitlines() if "Warning" not in l and not l.startswith("  ")]}
    for f in (out_file, cv_file):
        if os.path.exists(f):
            os.remove(f)
    json.dump(res, open(out_path, "w"), indent=1)


# ---------------------------------------------------------------------------------------------
# driver side
# ---------------------------------------------------------------------------------------------
def run_worker(kind, spec, out_path):
    spec_path = out_path + ".spec.json"
    json.dump(spec, open(spec_path, "w"))
    try:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", kind, spec_path, out_path],
                       check=True, env={**os.environ, "NUMBA_DISABLE_PERFORMANCE_WARNINGS": "1"})
    finally:
        os.remove(spec_path)
    print("wrote", out_path, flush=True)


def synth_counts(gen_pat, seed, style):
    """Small synthetic count tables that provoke the hard cases: zero k-mers, exact ties, skew."""
    import numpy as np

    rng = np.random.default_rng(seed)
    n = len(kmers_of(gen_pat))
    if style == "ties":          # few distinct values -> many exactly equal split sums
        U = rng.choice([0, 50, 100, 1000], size=n)
        M = np.minimum(U, rng.choice([0, 1, 2, 5], size=n))
    elif style == "sparse":      # most k-mers unseen
        U = rng.integers(0, 2000, size=n) * (rng.random(n) < 0.3)
        M = (rng.random(n) < 0.2) * rng.integers(0, 20, size=n)
    elif style == "big":         # genome-scale background (float32 ulp of the score ~ 0.1)
        U = rng.integers(10**5, 4 * 10**6, size=n)
        M = rng.binomial(U // 1000, 0.02 + 0.2 * rng.random(n))
    else:                        # "nb": negative-binomial background, position-dependent rate
        U = 1 + rng.negative_binomial(2, 2 / (2 + 3000.0), size=n)
        M = rng.binomial(U, np.minimum(0.5, 0.01 * np.exp(rng.normal(0, 0.7, size=n))))
    M = M.astype(np.int64)
    U = U.astype(np.int64)
    if M.sum() == 0:
        M[0] = 1
    if U.sum() == 0:
        U[0] = 1
    return [int(x) for x in M], [int(x) for x in U]


SINGLE_CASES = [
    # name, gen_pat, style, seed, alpha, penalty
    ("NN_nb", "NN", "nb", 1, 0.8, 3.0),
    ("NNN_nb", "NNN", "nb", 2, 1.0, 5.0),
    ("NNN_ties", "NNN", "ties", 3, 0.5, 2.0),
    ("NNN_sparse", "NNN", "sparse", 4, 10.0, 0.5),
    ("SWSW_ties", "SWSW", "ties", 5, 0.8, 1.0),
    ("BDHV_nb", "BDHV", "nb", 6, 0.8, 4.0),
    ("RNAVY_big", "RNAVY", "big", 7, 2.0, 6.0),
    ("NANN_big", "NANN", "big", 8, 0.8, 5.0),
    ("KNMNB_nb", "KNMNB", "nb", 9, 0.8, 3.0),
    ("NNNN_ties", "NNNN", "ties", 10, 0.8, 3.5),
    ("NNMNN_big", "NNMNN", "big", 11, 0.8, 5.0),
    ("A_nb", "A", "nb", 12, 0.8, 3.0),
    ("NTN_zeroalpha", "NTN", "sparse", 13, 0.0, 2.0),
]

CV_CASES = [
    # name, gen_pat, style, seed(data), alphas, penalties, nfolds, seed(cv), iterations
    ("NNN_nb", "NNN", "nb", 21, [0.5, 2.0], [2.0, 5.0], 3, 1, 1),
    ("NNN_ties", "NNN", "ties", 22, [0.8], [1.0, 3.0, 6.0], 2, 7, 1),
    ("NMNN_big", "NMNN", "big", 23, [0.5, 1.0, 10.0], [3.0, 6.0], 5, 1, 1),
    ("SNNB_nb", "SNNB", "nb", 24, [1.0], [4.0], 4, 3, 1),
    ("NNANN_big", "NNANN", "big", 25, [0.8, 4.0], [5.0], 2, 11, 1),
]


def beta_for(alpha, M, U):
    mu = sum(M) / (sum(M) + sum(U))
    return (alpha * (1.0 - mu)) / mu


def make_small():
    for name, gp, style, seed, alpha, pen in SINGLE_CASES:
        M, U = synth_counts(gp, seed, style)
        spec = {"gen_pat": gp, "M": M, "U": U, "alpha": alpha, "beta": beta_for(alpha, M, U) if alpha else 1.0,
                "penalty": pen}
        run_worker("single", spec, os.path.join(HERE, f"single_{name}.npz"))
    for name, gp, style, seed, alphas, pens, nf, cvseed, nit in CV_CASES:
        M, U = synth_counts(gp, seed, style)
        spec = {"gen_pat": gp, "M": M, "U": U, "alphas": alphas, "penalties": pens, "nfolds": nf, "seed": cvseed,
                "iterations": nit}
        run_worker("cv", spec, os.path.join(HERE, f"cv_{name}.npz"))


def copy_test_data():
    import shutil

    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    for f in ("mutated_5mers.txt", "background_5mers.txt", "mutated_7mers.txt", "background_7mers.txt"):
        shutil.copy(os.path.join(REF_DATA, f), os.path.join(HERE, "data", f))


def derive_inputs():
    """Input files for the other accepted formats, derived from the 5-mer test data:
    negative_5mers.txt (background - positive) and joint_5mers.txt (kmer positive background)."""
    d = os.path.join(HERE, "data")
    pos, bg = {}, {}
    for name, tab in (("mutated_5mers.txt", pos), ("background_5mers.txt", bg)):
        for line in open(os.path.join(d, name)):
            kmer, count = line.split()
            tab[kmer] = tab.get(kmer, 0) + int(count)
    with open(os.path.join(d, "negative_5mers.txt"), "w") as f:
        for kmer in sorted(bg):
            f.write(f"{kmer} {bg[kmer] - pos.get(kmer, 0)}\n")
    with open(os.path.join(d, "joint_5mers.txt"), "w") as f:
        for kmer in sorted(bg):
            f.write(f"{kmer}\t{pos.get(kmer, 0)}\t{bg[kmer]}\n")


def make_cli(which):
    d = os.path.join(HERE, "data")
    if which == "cli5_negative":   # --negative instead of --background
        argv = ["-p", f"{d}/mutated_5mers.txt", "-n", f"{d}/negative_5mers.txt", "-c", "5", "-a", "0.8"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_negative.json"))
    elif which == "cli5_joint":    # --joint_context_counts, long output
        argv = ["-j", f"{d}/joint_5mers.txt", "-c", "4", "-a", "2", "-l"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_joint.json"))
    elif which == "cli5_trim":     # 7-mer background collapsed around its centre to the 5-mers of the positive set
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_7mers.txt", "-c", "5", "-a", "0.8"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_trimmed_background.json"))
    elif which == "cli5_smallerk":  # --test_smaller_k: CV on 5-mers and on 3-mers, final DP on the better k
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "3", "6", "--test_smaller_k",
                "--seed", "2"]
        run_worker("cli", {"argv": argv, "zero_empty": True}, os.path.join(HERE, "cli_5mers_smaller_k.json"))
    elif which == "cli5_iter2":   # repeated CV: the reference's per-fold totals of iteration 2 include iteration 1
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "3", "6", "--nfolds", "3",
                "--iterations", "2", "--seed", "3"]
        run_worker("cli", {"argv": argv}, os.path.join(HERE, "cli_5mers_iterations2.json"))
    elif which == "cli5_verbose":   # --verbosity 2: per-level and per-fold progress lines
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "3", "5", "--seed", "1",
                "--verbosity", "2"]
        run_worker("cli", {"argv": argv}, os.path.join(HERE, "cli_5mers_verbose.json"))
    elif which == "cli5_allkmers":   # --score all_kmers: CV over pseudo counts of the one-rate-per-k-mer model
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "--score", "all_kmers",
                "-a", "0.5", "1", "10", "--nfolds", "3", "--seed", "2"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_all_kmers.json"))
    elif which == "cli5_greedy":     # the greedy (top-down) partition as the final fit
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "5", "-a", "0.8", "--greedy"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_greedy.json"))
    elif which == "cli5_greedycv":   # grid-search CV with the greedy partition, final fit with the optimal DP
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "3", "5", "-a", "0.5", "1",
                "--greedyCV", "--seed", "1"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_greedyCV.json"))
    elif which == "cli5_greedy_both":  # greedy CV (3 folds, 2 repeats) and greedy final fit
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "3", "6", "-a", "0.5", "2",
                "--greedy", "--nfolds", "3", "--iterations", "2", "--seed", "4"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_greedy_cv_and_fit.json"))
    elif which == "cli7_greedy":     # greedy final fit on the 7-mers
        argv = ["-p", f"{d}/mutated_7mers.txt", "-b", f"{d}/background_7mers.txt", "-c", "6", "-a", "10", "--greedy"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_7mers_greedy.json"))
    elif which == "cli5_scores":   # the information-criterion penalties (--score BIC)
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "--score", "BIC", "-a", "1"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_BIC.json"))
    elif which == "cli5":   # BASELINE config 1
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "3", "5", "7", "--seed", "1"]
        run_worker("cli", {"argv": argv}, os.path.join(HERE, "cli_cfg1_5mers.json"))
    elif which == "cli5_single":
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "5", "-a", "0.8"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_single.json"))
    elif which == "cli5_sp":  # super-pattern restricted, long output
        argv = ["-p", f"{d}/mutated_5mers.txt", "-b", f"{d}/background_5mers.txt", "-c", "4", "-s", "RNAYN", "-l"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_5mers_superpattern.json"))
    elif which == "cli7":  # BASELINE config 2 (about 26 min)
        argv = ["-p", f"{d}/mutated_7mers.txt", "-b", f"{d}/background_7mers.txt", "-c", "3", "5", "6",
                "-a", "0.5", "1", "10", "--nfolds", "5", "--seed", "1"]
        run_worker("cli", {"argv": argv}, os.path.join(HERE, "cli_cfg2_7mers.json"))
    elif which == "cli7_single":  # final DP of config 2 alone (about 1 min)
        argv = ["-p", f"{d}/mutated_7mers.txt", "-b", f"{d}/background_7mers.txt", "-c", "6", "-a", "10"]
        run_worker("cli", {"argv": argv, "cvfile": False}, os.path.join(HERE, "cli_7mers_single.json"))


def main():
    if len(sys.argv) >= 2 and sys.argv[1] == "--worker":
        kind, spec_path, out_path = sys.argv[2:5]
        {"single": worker_single, "cv": worker_cv, "cli": worker_cli}[kind](spec_path, out_path)
        return
    what = sys.argv[1:] or ["all"]
    copy_test_data()
    derive_inputs()
    for w in what:
        if w in ("all", "small"):
            make_small()
        if w == "all":
            for c in ("cli5", "cli5_single", "cli5_sp", "cli5_negative", "cli5_joint", "cli5_trim", "cli5_smallerk",
                      "cli5_iter2", "cli5_verbose", "cli5_allkmers", "cli5_greedy", "cli5_greedycv", "cli5_greedy_both", "cli7_greedy", "cli5_scores", "cli7_single", "cli7"):
                make_cli(c)
        elif w.startswith("cli"):
            make_cli(w)


if __name__ == "__main__":
    main()
