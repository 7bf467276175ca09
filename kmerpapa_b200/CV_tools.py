"""Fold sampling for cross validation, on the host.

Bit parity with the reference needs the exact stream of numpy's legacy RandomState, so this stays
numpy on the CPU (reference: src/kmerpapa/CV_tools.py:5-62).  The sampler deals every count of
every k-mer into nfolds held-out folds: k-mers sorted as strings, urn colours = all positive counts
followed by all negative counts, nfolds-1 sequential multivariate-hypergeometric draws of
total//nfolds balls each, the last fold takes what is left.
"""
import numpy as np


def draw_multivariate_hypergeometric(m, colors, prng):
    """One draw of m balls without replacement from an urn with colors[i] balls of colour i, as a
    chain of univariate hypergeometric draws (same call sequence as the reference's `sample`)."""
    ncol = len(colors)
    tail = np.cumsum(colors[::-1])[::-1]  # tail[i] = balls of colour >= i
    picked = np.zeros(ncol, dtype=colors.dtype)
    for i in range(ncol - 1):
        if m < 1:
            break
        picked[i] = prng.hypergeometric(colors[i], tail[i + 1], m)
        m -= picked[i]
    picked[-1] = m
    return picked


def iter_fold_counts(kmers, pos, neg, nfolds, prng, itype=np.uint64):
    """The same sampler as sample_fold_counts, one fold at a time: yields (f, M, U) with the held-out counts of
    fold f in the order of `kmers`.  Fold f is final as soon as it has been drawn, so a consumer can start the
    jobs of fold f while fold f + 1 is still being drawn."""
    n = len(kmers)
    order = sorted(range(n), key=kmers.__getitem__)
    inv = np.asarray(order)
    urn = np.empty(2 * n, dtype=itype)
    urn[:n] = np.asarray(pos, dtype=itype)[inv]
    urn[n:] = np.asarray(neg, dtype=itype)[inv]
    per_fold = urn.sum() // nfolds
    for f in range(nfolds):
        if f < nfolds - 1:
            got = draw_multivariate_hypergeometric(per_fold, urn, prng)
            urn -= got
        else:
            got = urn
        M = np.empty(n, dtype=itype)
        U = np.empty(n, dtype=itype)
        M[inv] = got[:n]
        U[inv] = got[n:]
        yield f, M, U


def sample_fold_counts(kmers, pos, neg, nfolds, prng, itype=np.uint64, sort=True):
    """Held-out counts per fold.

    kmers: list of k-mer strings; pos/neg: their counts (same order).
    Returns (Mf, Uf), arrays [len(kmers), nfolds] in the order of `kmers`.
    sort=True: urn colours in sorted k-mer order (the pattern-partition CV, CV_tools.py:42-43 of the reference);
    sort=False: in the given order (the all-k-mers CV enumerates `matches(gen_pat)`, :75).
    """
    n = len(kmers)
    order = sorted(range(n), key=kmers.__getitem__) if sort else list(range(n))
    urn = np.empty(2 * n, dtype=itype)
    for r, i in enumerate(order):
        urn[r] = pos[i]
        urn[n + r] = neg[i]
    per_fold = urn.sum() // nfolds
    draws = np.empty((2 * n, nfolds), dtype=itype)
    for f in range(nfolds - 1):
        got = draw_multivariate_hypergeometric(per_fold, urn, prng)
        draws[:, f] = got
        urn -= got
    draws[:, nfolds - 1] = urn
    Mf = np.empty((n, nfolds), dtype=itype)
    Uf = np.empty((n, nfolds), dtype=itype)
    inv = np.asarray(order)
    Mf[inv] = draws[:n]
    Uf[inv] = draws[n:]
    return Mf, Uf


def make_all_folds(kmer_table, n_folds, n_repeats, prng):
    """Folds of a [n_kmers, 2] count table for the greedy estimator's CV (reference CV_tools.py:123-147): the urn
    colours are the table's entries in row-major order (positive, negative of k-mer 0, then k-mer 1, ...), every
    repeat deals the whole table again.  Returns [n_repeats, n_folds, n_kmers, 2]."""
    itype = kmer_table.dtype
    shape = kmer_table.shape
    folds = np.zeros((n_repeats, n_folds) + shape, dtype=itype)
    per_fold = kmer_table.sum() // n_folds
    for i in range(n_repeats):
        urn = np.copy(kmer_table).reshape(-1)
        for j in range(n_folds - 1):
            got = draw_multivariate_hypergeometric(per_fold, urn, prng)
            urn -= got
            folds[i][j] = got.reshape(shape)
        folds[i][n_folds - 1] = urn.reshape(shape)
    return folds
