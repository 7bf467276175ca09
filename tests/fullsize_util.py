"""Index helpers of the full-size parity tests (no GPU needed): where a sub-lattice sits inside a larger lattice."""
import numpy as np


def sublattice_patnums(gen_pat, sub):
    """Dense numbers (in gen_pat's numbering) of every sub-pattern of `sub`, in the order of sub's own numbering."""
    from kmerpapa_b200 import iupac

    n = iupac.pattern_max(sub)
    j = np.arange(n, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    w_full = 1
    for cf, cs in zip(gen_pat, sub):
        rf, rs = len(iupac.PERM[cf]), len(iupac.PERM[cs])
        lut = np.array([iupac.PERM[cf].index(x) for x in iupac.PERM[cs]], dtype=np.uint64)
        if rs == 1:
            out += np.uint64(int(lut[0]) * w_full)
        else:
            out += lut[(j % np.uint64(rs)).astype(np.int64)] * np.uint64(w_full)
            j //= np.uint64(rs)
        w_full *= rf
    return out


def sub_kmer_select(gen_pat, sub):
    """Indices (k-mer index order of gen_pat) of the k-mers of `sub`, in sub's k-mer index order."""
    from kmerpapa_b200 import iupac

    n = len(iupac.matches(sub))
    j = np.arange(n, dtype=np.int64)
    out = np.zeros(n, dtype=np.int64)
    w_full = 1
    for cf, cs in zip(gen_pat, sub):
        bf, bs = iupac.CODE[cf], iupac.CODE[cs]
        lut = np.array([bf.index(x) for x in bs], dtype=np.int64)
        out += lut[j % len(bs)] * w_full
        j //= len(bs)
        w_full *= len(bf)
    return out
