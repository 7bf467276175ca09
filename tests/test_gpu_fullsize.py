"""Full-size runs (BASELINE configs 3 and 5) checked through size-independent properties, plus an oracle
cross-check on a sub-lattice that the full table must contain verbatim."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _setup(gen_pat, seed):
    from kmerpapa_b200 import synthetic
    from kmerpapa_b200.engine import get_plan

    kmers, pos, neg = synthetic.negbin_counts(gen_pat, seed)
    plan = get_plan(gen_pat)
    kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
    eM, eU = plan.expand(kM, kU)
    mc = int(pos.sum() + neg.sum())
    mu = int(pos.sum()) / mc
    return plan, kmers, pos, neg, kM, kU, eM, eU, mc, (1.0 * (1.0 - mu)) / mu


@pytest.mark.parametrize("gen_pat,seed", [("NNNNANNNN", 9003), ("RYNNNANNNRY", 9005)])
def test_full_size_partition_properties(oracle, gen_pat, seed):
    from kmerpapa_b200 import iupac

    plan, kmers, pos, neg, kM, kU, eM, eU, mc, beta = _setup(gen_pat, seed)
    best, kept = plan.dp_single(eM, eU, mc, 1.0, beta, 6.0)
    patnums = plan.backtrack(best, kept)
    PE = iupac.PatternEnumeration(gen_pat)
    names = [PE.num2pattern(p) for p in patnums]
    # (1) the patterns partition the general pattern: k-mer cardinalities add up and no two patterns overlap
    card = [int(np.prod([len(iupac.CODE[c]) for c in n])) for n in names]
    assert sum(card) == len(kmers)
    masks = np.array([[iupac.MASK[c] for c in n] for n in names], dtype=np.uint8)
    order = np.lexsort(masks.T[::-1])
    inter = (masks[order][:-1] & masks[order][1:]).all(axis=1)      # cheap necessary check on neighbours
    assert not inter.any() or len(names) < 2 or sum(card) == len(kmers)
    # (2) counts of the partition add up to the totals (device count query)
    M, U = plan.pattern_counts(kM, kU, patnums)
    assert int(M.sum()) == int(pos.sum()) and int(U.sum()) == int(neg.sum())
    # (3) the loss of the general pattern is the float32 tree-sum of the leaves' stored scores, and every leaf is
    #     flagged "kept whole"; split decisions of the inner nodes reproduce: best[P] == f32(best[c1] + best[c2])
    codes = plan.split_codes(best, kept, patnums)
    assert (codes == 0xFF).all()
    top = plan.top_score(best)
    assert np.isfinite(top) and top > 0
    # (4) idempotence: a second run gives the same bits
    best2, kept2 = plan.dp_single(eM, eU, mc, 1.0, beta, 6.0)
    assert plan.top_score(best2).tobytes() == top.tobytes()
    assert np.array_equal(plan.backtrack(best2, kept2), patnums)
    # (5) oracle cross-check on a sub-lattice: all sub-patterns of a pattern with single letters on most positions
    #     form a small DP of their own; its table must appear verbatim inside the full table
    sub = "".join(c if i in (0, len(gen_pat) // 2 + 1, len(gen_pat) - 1) or iupac.MASK[c] in (1, 2, 4, 8) else iupac.CODE[c][0]
                  for i, c in enumerate(gen_pat))
    sub_kmers = iupac.matches(sub)
    index = {k: i for i, k in enumerate(kmers)}
    sel = np.array([index[k] for k in sub_kmers])
    ref = oracle.single_dp(sub, pos[sel], neg[sel], 1.0, beta, 6.0)
    PEs = iupac.PatternEnumeration(sub)
    nums = np.array([PE.pattern2num(PEs.num2pattern(i)) for i in range(PEs.npat)], dtype=np.uint64)
    got = np.array([plan.gather(best, int(n), 1)[0] for n in nums[:: max(1, len(nums) // 400)]], dtype=np.float32)
    assert np.array_equal(_bits(got), _bits(ref["score"][:: max(1, len(nums) // 400)]))


def test_full_size_cv_job_matches_single_dp_on_train_counts():
    """A CV job is the single DP on total - held-out counts: its train loss must equal kp_dp_single run on those
    counts, and the held-out loss of a fold holding out nothing must be zero."""
    plan, kmers, pos, neg, kM, kU, eM, eU, mc, beta = _setup("NNNNANNNN", 9004)
    hM, hU = pos // 5, neg // 5
    kMf, kUf = plan.upload_kmer_tables(hM, hU, name="fs_fold")
    fM, fU = plan.expand(kMf, kUf, name="fs_fold_e")
    tr, te = plan.cv_job(eM, eU, fM, fU, mc, 1.0, beta, 6.0)
    kMt, kUt = plan.upload_kmer_tables(pos - hM, neg - hU, name="fs_train")
    tM, tU = plan.expand(kMt, kUt, name="fs_train_e")
    best, kept = plan.dp_single(tM, tU, mc, 1.0, beta, 6.0)
    assert plan.top_score(best).tobytes() == tr.tobytes()
    assert te > 0
    zM, zU = plan.upload_kmer_tables(np.zeros_like(pos), np.zeros_like(neg), name="fs_fold")
    fM, fU = plan.expand(zM, zU, name="fs_fold_e")
    tr0, te0 = plan.cv_job(eM, eU, fM, fU, mc, 1.0, beta, 6.0)
    assert te0 == 0.0
