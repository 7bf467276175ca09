"""The `kmerpapa` command line, unchanged on the outside, GPU underneath.

Flags, defaults, stderr messages, the partition file and the CV file are those of the reference's
src/kmerpapa/cli.py (get_parser :16-115, main :118-318).  What differs is below the two estimator
calls (cli.py:232 and :279): they go to kmerpapa_b200.algorithms (CUDA), and the per-pattern counts
of the output rows come from the device k-mer tables (kp_pattern_counts) instead of the reference's
Python enumeration (pattern_utils.get_M_U, cli.py:281-283).

`--score all_kmers` goes to kmerpapa_b200.algorithms.all_kmers_CV, `--greedy` / `--greedyCV` to
kmerpapa_b200.algorithms.greedy_penalty_plus_pseudo.  --BayesOpt (scikit-optimize) is accepted by the parser like in
the reference but stops with a clear error.
"""
import argparse
import sys
from math import log

import numpy as np

from . import __version__, iupac
from .io_utils import downsize_contextD_device as downsize_contextD, read_input
from .score_utils import beta_from_totals, get_loss


def get_parser():
    p = argparse.ArgumentParser(prog="kmerpapa", description="Finds optimal k-mer pattern partition in fx. mutation data")
    p.add_argument("-p", "--positive", type=argparse.FileType("r"), help="File with k-mer counts in positive set")
    p.add_argument("-n", "--negative", type=argparse.FileType("r"),
                   help="File with k-mer counts in negative set. Longer k-mers are collapsed around their centre to the "
                        "length of the positive k-mers.")
    p.add_argument("-b", "--background", type=argparse.FileType("r"),
                   help="File with k-mer counts in background set (positive and negative regions together). Longer "
                        "k-mers are collapsed around their centre to the length of the positive k-mers.")
    p.add_argument("-j", "--joint_context_counts", type=argparse.FileType("r"),
                   help="File with three columns: k-mer, positive count, background count. Replaces -p/-b.")
    p.add_argument("-o", "--output", type=argparse.FileType("w"), default="-", metavar="PATH",
                   help="Output file (default: standard output)")
    p.add_argument("-f", "--CVfile", type=argparse.FileType("w"),
                   help="File with the held-out likelihood of every (pseudo count, penalty) pair tried.")
    p.add_argument("--verbosity", type=int, default=1, help="0: silent, 1: default, 2: verbose (stderr)")
    p.add_argument("--CV_only", action="store_true", help="Only run cross validation, no final fit.")
    p.add_argument("--greedy", action="store_true", help="Use the greedy (top-down) partition instead of the optimal one.")
    p.add_argument("--BayesOpt", action="store_true", help="(reference heuristic; not part of this build)")
    p.add_argument("--greedyCV", action="store_true", help="Cross validate with the greedy partition, then fit the optimal one.")
    p.add_argument("-l", "--long_output", action="store_true", help="Print one row per k-mer instead of one per pattern.")
    p.add_argument("-s", "--super_pattern", type=str,
                   help="Only k-mers matching this IUPAC pattern are used, e.g. NNANN when the positive file only "
                        "holds mutations of A.")
    p.add_argument("--score", type=str, default="penalty_and_pseudo",
                   choices=["penalty_and_pseudo", "all_kmers", "BIC", "AIC", "HQ", "LL"],
                   help='Score function (default "penalty_and_pseudo").')
    p.add_argument("-N", "--nfolds", "--n_folds", dest="nfolds", type=int, metavar="N",
                   help="Cross validation with N folds (default 2 when several pseudo counts / penalties are given).")
    p.add_argument("-i", "--iterations", type=int, default=1, metavar="i", help="Repeat cross validation i times")
    p.add_argument("-a", "--pseudo_counts", type=float, metavar="a", nargs="+", default=[0.8],
                   help="Pseudo count (alpha) values to try")
    p.add_argument("-c", "--penalty_values", type=float, metavar="c", nargs="+",
                   help="Penalty values to try (default for the standard score: log(#k-mers))")
    p.add_argument("--test_smaller_k", action="store_true",
                   help="Also cross-validate every smaller odd k and keep the best.")
    p.add_argument("--seed", type=int, help="seed for numpy.random")
    p.add_argument("-V", "--version", action="store_true", help="Print version number and return")
    return p


def _default_penalties(args, n_mut, n_kmers):
    if args.score == "BIC":
        return [log(n_mut)]
    if args.score == "AIC":
        return [2.0]
    if args.score == "HQ":
        return [log(log(n_mut))]
    if args.score == "LL":
        return [0.0]
    if args.score == "penalty_and_pseudo" and not args.BayesOpt:
        pen = [log(n_kmers)]
        if args.verbosity > 0:
            print(f"penalty values not set. Using {pen[0]}", file=sys.stderr)
        return pen
    return None


def _pattern_counts(gen_pat, contextD, names):
    """(M, U) of each output pattern, from the device k-mer tables."""
    from .algorithms.bottum_up_array_w_numba import kmer_arrays
    from .engine import get_plan

    plan = get_plan(gen_pat, lite=True)
    codes, pos, neg = kmer_arrays(contextD)
    kM, kU = plan.pack_counts(codes, pos, neg, name="out_k")
    PE = iupac.PatternEnumeration(gen_pat)
    M, U = plan.pattern_counts(kM, kU, np.array([PE.pattern2num(p) for p in names], dtype=np.uint64))
    return [(int(m), int(u)) for m, u in zip(M, U)]


def _init_distributed(args=None):
    """Under torchrun (WORLD_SIZE > 1) every process takes one GPU and joins an NCCL group: the CV grid is then
    sharded by job across the GPUs.  Returns this process's rank.

    Every rank samples the folds itself (same RandomState stream), so all ranks need the SAME seed: without --seed the
    reference seeds from OS entropy (RandomState(None)); here rank 0 draws that seed and broadcasts it, otherwise the
    ranks would cross-validate on different fold samplings and could enter the final (collective) fit with different
    parameters."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    share_seed(args)
    return dist.get_rank()


def share_seed(args):
    """All ranks of the process group leave with the same args.seed (rank 0's; drawn from OS entropy when unset)."""
    import torch.distributed as dist

    if args is None or not dist.is_initialized() or dist.get_world_size() <= 1:
        return
    box = [None]
    if dist.get_rank() == 0:
        seed = args.seed
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])   # 32 bits of OS entropy: a valid RandomState seed
        box[0] = seed
    dist.broadcast_object_list(box, src=0)
    args.seed = box[0]


def main(args=None):
    """Runs the program; returns the exit code (0 also on input errors, like the reference)."""
    parser = get_parser()
    args = parser.parse_args(args=args)
    rank = _init_distributed(args)
    if rank != 0:   # only rank 0 reports and writes files; the others just run their share of the CV jobs
        import os

        args.verbosity = 0
        args.output = open(os.devnull, "w")
        if args.CVfile is not None:
            args.CVfile = open(os.devnull, "w")
    if args.version:
        print("version:", __version__)
        print()
        return 0
    super_pattern = args.super_pattern
    try:
        contextD, n_unmut, n_mut = read_input(args, super_pattern)
    except Exception as e:  # the reference prints the help and the message, and exits 0
        parser.print_help()
        print("=" * 80, file=sys.stderr)
        print("input error:", file=sys.stderr)
        print(e, file=sys.stderr)
        print("=" * 80, file=sys.stderr)
        return 0
    if args.verbosity > 0:
        print(f"Input data read. {n_mut} positive k-mers and {n_unmut} negative k-mers", file=sys.stderr)
    if args.BayesOpt:
        raise SystemExit("--BayesOpt needs scikit-optimize (skopt), which is not part of this build; "
                         "use the grid search (--greedyCV / --penalty_values / --pseudo_counts)")
    if args.penalty_values is not None:
        assert args.score == "penalty_and_pseudo", \
            f"you cannot specify penalty values when using the {args.score} score function"
    else:
        args.penalty_values = _default_penalties(args, n_mut, len(contextD))

    gen_pat = iupac.lca_pattern(list(contextD.keys()))
    if args.super_pattern is not None:
        assert gen_pat == args.super_pattern
    for kmer in iupac.matches(gen_pat):
        if kmer not in contextD:
            contextD[kmer] = (0, 0)
    if args.verbosity > 0:
        print(f"General pattern: {gen_pat}", file=sys.stderr)
    if args.CVfile is not None:
        print("k alpha P LL_test", file=args.CVfile)

    from .algorithms import (all_kmers_CV, bottum_up_array_penalty_plus_pseudo_CV, bottum_up_array_w_numba,
                             greedy_penalty_plus_pseudo)

    best_alpha = best_penalty = best_k = None
    ks = range(len(gen_pat), 1, -2) if args.test_smaller_k else [len(gen_pat)]
    this_contextD, this_gen_pat = contextD, gen_pat
    best_score = 1e100
    if args.nfolds is None and (len(ks) > 1 or len(args.pseudo_counts) > 1 or len(args.penalty_values) > 1 or args.CV_only):
        args.nfolds = 2
    if args.nfolds is not None and args.nfolds > 1:
        for k in ks:
            if args.verbosity > 0:
                print(f"Running {args.nfolds}-fold cross validation on {k}-mers", file=sys.stderr)
            if k != len(this_gen_pat):
                this_contextD, this_gen_pat = downsize_contextD(this_contextD, this_gen_pat, k)
            if args.greedy or args.greedyCV:   # grid search with the greedy partition (cli.py:217-225 of the reference)
                assert args.score != "all_kmers", "greedy option cannot be used wil all-kmers"
                CV = greedy_penalty_plus_pseudo.GridSearchCV(gen_pat, contextD, args.penalty_values, args.pseudo_counts,
                                                             args.nfolds, args.iterations, args.seed)
                this_alpha, this_penalty, test_score = CV.get_best_a_c()
            elif args.score == "all_kmers":
                this_alpha, test_score = all_kmers_CV.all_kmers(this_gen_pat, this_contextD, args.pseudo_counts, args,
                                                                n_mut, n_unmut)
                this_penalty = None
            else:
                this_alpha, this_penalty, test_score = bottum_up_array_penalty_plus_pseudo_CV.pattern_partition_bottom_up(
                    this_gen_pat, this_contextD, args.pseudo_counts, args, n_mut, n_unmut, args.penalty_values)
            with np.errstate(over="ignore"):   # np.float32 against the 1e100 start value, as in the reference
                better = test_score < best_score
            if better:
                best_score, best_k, best_alpha, best_penalty = test_score, k, this_alpha, this_penalty
        if args.verbosity > 0:
            print(f"CV DONE. best_k={best_k}, best_alpha={best_alpha}, best_penalty={best_penalty}, "
                  f"best_test_LL={best_score}", file=sys.stderr)
    if args.CVfile is not None:
        args.CVfile.close()
    if args.CV_only:
        return 0

    if best_alpha is None:
        assert len(args.pseudo_counts) == 1
        best_alpha = args.pseudo_counts[0]
    if args.score != "all_kmers" and best_penalty is None:
        assert len(args.penalty_values) == 1
        best_penalty = args.penalty_values[0]
    if best_k is None:
        best_k = len(gen_pat)
    if best_k != len(gen_pat):
        contextD, gen_pat = downsize_contextD(contextD, gen_pat, best_k)
    best_beta = beta_from_totals(best_alpha, n_mut, n_unmut)
    if args.verbosity > 0:
        print(f"Training on whole data set with k={best_k} alpha={best_alpha} penalty={best_penalty}", file=sys.stderr)

    if args.score == "all_kmers":   # every k-mer is its own pattern (cli.py:267-272 of the reference)
        best_score, M, U, names = 0, n_mut, n_unmut, list(iupac.matches(gen_pat))
        counts = [tuple(contextD[k][:2]) for k in names]
    elif args.greedy:
        best_score, M, U, names = greedy_penalty_plus_pseudo.greedy_partition(gen_pat, contextD, best_alpha, best_beta,
                                                                              best_penalty, args)
        counts = _pattern_counts(gen_pat, contextD, names)
    else:
        best_score, M, U, names = bottum_up_array_w_numba.pattern_partition_bottom_up(
            gen_pat, contextD, best_alpha, best_beta, best_penalty, args, n_mut, n_unmut)
        counts = _pattern_counts(gen_pat, contextD, names)
    assert M == n_mut
    assert U == n_unmut
    assert n_mut == sum(x[0] for x in counts)
    assert n_unmut == sum(x[1] for x in counts)

    if args.verbosity > 0:
        print(f"Optimal k-mer pattern partition contains {len(names)} patterns.", file=sys.stderr)
        print(f"loss={best_score}", file=sys.stderr)
        print(f"LL={get_loss(counts, best_alpha, best_beta)}", file=sys.stderr)

    if args.long_output:
        print("context", "c_neg", "c_pos", "c_rate", "pattern", "p_neg", "p_pos", "p_rate", file=args.output)
    else:
        print("pattern", "p_neg", "p_pos", "p_rate", file=args.output)
    for pat, (Mp, Up) in zip(names, counts):
        p = (Mp + best_alpha) / (Mp + Up + best_alpha + best_beta)
        if args.long_output:
            for kmer in iupac.matches(pat):
                nm, ns = contextD[kmer]
                print(kmer, ns, nm, float(nm) / (nm + ns), pat, Up, Mp, p, file=args.output)
        else:
            print(pat, Up, Mp, p, file=args.output)
    args.output.flush()
    return 0
