"""Reading k-mer count files into {kmer: (n_pos, n_neg)} tables.

Host-side mirror of the reference's src/kmerpapa/io_utils.py (read_dict :82-136,
read_postive_and_other :139-184, read_joint_kmer_counts :3-47, read_input :187-217,
downsize_contextD :50-79).  Same accepted inputs, same filters, same assertion behaviour; the
resulting table is handed to the GPU through kmerpapa_b200.algorithms.* (K1 packs it).
"""
from . import iupac


def _to_count(text):
    try:
        return int(text)
    except ValueError:
        return int(float(text))


_NO_ACGT = {ord(c): None for c in "ACGT"}


def _is_plain_kmer(kmer):
    return not kmer.translate(_NO_ACGT)   # nothing is left once A, C, G, T are deleted


def _centre_window(have, want):
    """Slice that keeps the centred `want` positions of a k-mer of length `have`."""
    start = have // 2 - want // 2
    return start, start + want


def read_dict(f, super_pattern, length=None):
    """`kmer count` lines -> ({kmer: count}, total).  k-mers with non-ACGT letters are skipped,
    longer k-mers are trimmed around the centre to `length` (or the super-pattern's length),
    k-mers outside the super-pattern are dropped, repeated k-mers add up."""
    if length is None and super_pattern is not None:
        length = len(super_pattern)
    table, total, window = {}, 0, None
    for line in f:
        kmer, count = line.split()
        if not _is_plain_kmer(kmer):
            continue
        count = _to_count(count)
        assert count >= 0, f"negative counts are not allowed, bad line:\n{line.strip()}"
        if window is None:
            if length is not None and length != len(kmer):
                assert len(kmer) > length
                window = _centre_window(len(kmer), length)
            else:
                window = (0, len(kmer))
        kmer = kmer[window[0]:window[1]]
        if super_pattern is not None:
            assert len(super_pattern) == len(kmer)
            if not iupac.contains(super_pattern, kmer):
                continue
        total += count
        table[kmer] = table.get(kmer, 0) + count
    return table, total


def read_postive_and_other(fpos, fother, super_pattern, n_scale=1, background=True):
    """Positive counts plus either negative counts or background (= positive + negative) counts."""
    pos, all_pos = read_dict(fpos, super_pattern)
    other, all_other = read_dict(fother, super_pattern, length=len(next(iter(pos.keys()))))
    table = {}
    for kmer in set([*pos.keys(), *other.keys()]):
        n_pos = pos.get(kmer, 0)
        n_other = n_scale * other[kmer] if kmer in other else 0
        if background:
            assert n_other >= n_pos, """
                background counts should be larger than the positive counts
                so that a negative set can be created by subtraction the positive count
                from the background count. Problematic k-mer: {context}
                """
            n_other -= n_pos
        table[kmer] = (n_pos, n_other)
    if background:
        all_other -= all_pos
    return table, all_other, all_pos


def read_joint_kmer_counts(f, super_pattern, n_scale=1):
    """`kmer count_pos count_background` lines."""
    table, n_sites, n_pos_total = {}, 0, 0
    for line in f:
        kmer, n_pos, n_bg = line.split()
        if not _is_plain_kmer(kmer):
            continue
        n_bg, n_pos = _to_count(n_bg), _to_count(n_pos)
        assert n_scale * n_bg - n_pos >= 0, f"""
            background counts should be larger than the positive counts
            so that a negative set can be created by subtraction the positive count
            from the background count. Problematic kmer: {kmer}"""
        if super_pattern is not None and not iupac.contains(super_pattern, kmer):
            continue
        n_sites += n_scale * n_bg
        n_pos_total += n_pos
        table[kmer] = (n_pos, n_scale * n_bg - n_pos)
    f.close()
    return table, n_sites - n_pos_total, n_pos_total


def read_input(args, super_pattern):
    """Returns (contextD, n_unmut, n_mut) from whichever input flags were given."""
    assert (args.positive is None) != (args.joint_context_counts is None), """
        Either the --positive option or the --join_context_counts option (but not both)
        must be used to provide input data.
        """
    if args.positive is not None:
        assert (args.negative is None) != (args.background is None), """
            If the --joint_context_counts option is not used then either the --negative or the
            --background option (but not both) must be used.
            """
        if args.negative is not None:
            return read_postive_and_other(args.positive, args.negative, super_pattern, n_scale=1, background=False)
        return read_postive_and_other(args.positive, args.background, super_pattern, n_scale=1, background=True)
    return read_joint_kmer_counts(args.joint_context_counts, super_pattern, n_scale=1)


def downsize_contextD(table, general_pattern, length):
    """Collapse a k-mer table to the centred `length`-mers (used by --test_smaller_k)."""
    out, window = {}, None
    for kmer, counts in table.items():
        if window is None:
            assert length is not None
            assert len(kmer) > length, f"k-mer:{kmer} cannot be reduced to length {length}"
            window = _centre_window(len(kmer), length)
        short = kmer[window[0]:window[1]]
        acc = out.setdefault(short, [0] * len(counts))
        for i, c in enumerate(counts):
            acc[i] += c
    return out, general_pattern[window[0]:window[1]]


def downsize_contextD_device(table, general_pattern, length):
    """downsize_contextD with the sums formed on the GPU.  Collapsing a k-mer to its centred
    `length`-mer is a shift of its packed 4-bit code, and summing the counts of equal short k-mers is what the packing
    kernel K1 (kp_pack_counts: atomic adds into the dense k-mer table) already does for duplicate input lines, so the
    short table is K1 on the shifted codes.  Same return value as downsize_contextD, keys in the same (first seen) order."""
    import numpy as np

    from . import iupac
    from .engine import get_plan

    kmers = list(table.keys())
    vals = list(table.values())
    if not kmers or any(len(v) != 2 for v in vals) or len(kmers[0]) > 16:
        return downsize_contextD(table, general_pattern, length)
    assert length is not None
    assert len(kmers[0]) > length, f"k-mer:{kmers[0]} cannot be reduced to length {length}"
    lo, hi = _centre_window(len(kmers[0]), length)
    short_gen = general_pattern[lo:hi]
    codes = (iupac.kmer_codes(kmers) >> np.uint64(4 * lo)) & np.uint64((1 << (4 * length)) - 1)
    pos = np.array([v[0] for v in vals], dtype=np.int64)
    neg = np.array([v[1] for v in vals], dtype=np.int64)
    plan = get_plan(short_gen, lite=True)
    kM, kU = plan.pack_counts(codes, pos, neg, name="downsize_k")
    M, U = kM[: plan.nkmer].cpu().numpy(), kU[: plan.nkmer].cpu().numpy()
    index = {km: i for i, km in enumerate(iupac.matches(short_gen))}
    out = {}
    for kmer in kmers:
        short = kmer[lo:hi]
        if short not in out:
            i = index[short]
            out[short] = [int(M[i]), int(U[i])]
    return out, short_gen
