"""Cross validation of the all-k-mers model (every k-mer its own rate): picks the pseudo count.

Drop-in for the reference's src/kmerpapa/algorithms/all_kmers_CV.py (`all_kmers`, :15-63; called from cli.py:229 for
`--score all_kmers`): same name, arguments, stderr lines and return value.  The fold sampler replays the reference's
numpy RandomState stream (CV_tools.make_all_folds_contextD_kmers, :65-95: k-mers in enumeration order, not sorted);
the per-(k-mer, fold) float64 terms (scipy xlogy / xlog1py formulas, :8-13) are evaluated on the GPU by
kp_kmer_fold_terms; their sums run over the k-mers in enumeration order, sequentially, like the reference's `+=`.
"""
import ctypes
import sys

import numpy as np

from .. import CV_tools, _native, iupac
from .._native import check
from ..score_utils import get_betas


def fold_terms(Mtr, Utr, Mte, Ute, betas, alpha, device=None):
    """float64 (train, test) terms of every (k-mer, fold): arrays [n_kmers, n_folds]."""
    from ..engine import _torch

    torch = _torch()
    device = torch.cuda.current_device() if device is None else int(device)
    shape = Mtr.shape
    args = [np.ascontiguousarray(x, dtype=np.int64).ravel() for x in (Mtr, Utr, Mte, Ute)]
    beta = np.ascontiguousarray(np.broadcast_to(np.asarray(betas, dtype=np.float64), shape)).ravel()
    train = np.empty(beta.size, dtype=np.float64)
    test = np.empty(beta.size, dtype=np.float64)
    check(_native.lib().kp_kmer_fold_terms(device, *[a.ctypes.data_as(ctypes.c_void_p) for a in args],
                                           beta.ctypes.data_as(ctypes.c_void_p), beta.size, float(alpha),
                                           train.ctypes.data_as(ctypes.c_void_p), test.ctypes.data_as(ctypes.c_void_p)),
          "kp_kmer_fold_terms")
    return train.reshape(shape), test.reshape(shape)


def all_kmers(gen_pat, contextD, alphas, args, nmut, nunmut, index_mut=0):
    """Returns (best_alpha, best_test_loss)."""
    nf, nit = args.nfolds, args.iterations
    kmers = iupac.matches(gen_pat)
    pos = np.array([contextD[k][0] for k in kmers], dtype=np.uint64)
    neg = np.array([contextD[k][1] for k in kmers], dtype=np.uint64)
    test_loss = {a_i: [] for a_i in range(len(alphas))}
    train_loss = {a_i: [] for a_i in range(len(alphas))}
    prng = np.random.RandomState(args.seed)
    for _ in range(nit):
        Mf, Uf = CV_tools.sample_fold_counts(kmers, pos, neg, nf, prng, sort=False)
        M_sum_test, U_sum_test = Mf.sum(axis=0), Uf.sum(axis=0)
        M_sum_train = sum(M_sum_test) - M_sum_test
        U_sum_train = sum(U_sum_test) - U_sum_test
        Mtr = Mf.sum(axis=1, keepdims=True) - Mf
        Utr = Uf.sum(axis=1, keepdims=True) - Uf
        for a_i, alpha in enumerate(alphas):
            betas = get_betas(alpha, M_sum_train, U_sum_train)
            tr, te = fold_terms(Mtr, Utr, Mf, Uf, betas, alpha)
            # sum_train += terms[k-mer] for the k-mers in order: a sequential float64 sum per fold
            sum_train = np.cumsum(np.vstack([np.zeros((1, nf)), tr]), axis=0)[-1]
            sum_test = np.cumsum(np.vstack([np.zeros((1, nf)), te]), axis=0)[-1]
            train_loss[a_i].extend(list(sum_train))
            test_loss[a_i].extend(list(sum_test))
    best_test_loss, best_alpha = 1e100, None
    for a_i, alpha in enumerate(alphas):
        test = sum(test_loss[a_i]) / nit
        if args.verbosity > 0:
            print(f"alpha={alpha} test_loss={test}", file=sys.stderr)
        if test < best_test_loss:
            best_alpha, best_test_loss = alpha, test
    return best_alpha, best_test_loss
