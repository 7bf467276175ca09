"""The CPU oracle against the reference's own outputs (tests/golden, made by make_golden.py).

These pin the oracle: kpo_log == glibc log bit for bit, cephes log1p == scipy, and full
float32 score tables / counts / partitions / CV rows == the unmodified numba reference.
"""
import numpy as np
import pytest

from conftest import golden_files


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


def test_log_matches_system_libm(oracle):
    """kpo_log restates glibc 2.39 __log_fma; every input must give the same bits as log()."""
    rng = np.random.default_rng(1)
    n = 2_000_000
    sets = [
        rng.random(n),                                   # rates p
        1.0 - rng.random(n) * 0.07,                      # 1-p for small p (near-1 polynomial branch)
        1.0 + (rng.random(n) - 0.5) * 0.2,               # both sides of the branch boundaries
        np.exp(rng.uniform(-700, 700, n)),               # whole exponent range
        rng.random(n) * 1e-300,                          # towards subnormal
        np.abs(rng.integers(0, 2**63 - 1, n, dtype=np.int64).view(np.float64)),  # random bit patterns
        np.array([0.0, -0.0, 1.0, np.inf, -1.0, np.nan, 5e-324, 2.2250738585072014e-308, 0.9375, 1.064697265625]),
    ]
    L = oracle.lib()
    for x in sets:
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert L.kpo_log_mismatches(x.ctypes.data, len(x)) == 0


def test_log1p_matches_scipy(oracle):
    import scipy.special as sp

    rng = np.random.default_rng(2)
    L = oracle.lib()
    for x in (-rng.random(500_000) * 0.5, -rng.random(500_000), -np.exp(rng.uniform(-40, 0, 500_000)),
              rng.random(100_000) * 3):
        x = np.ascontiguousarray(x)
        y = np.empty_like(x)
        L.kpo_log1p_array(x.ctypes.data, y.ctypes.data, len(x))
        assert np.array_equal(_bits(y), _bits(sp.log1p(x)))
        assert np.array_equal(_bits(7.0 * y), _bits(sp.xlog1py(np.full_like(x, 7.0), x)))


def test_index_bijection_and_levels(oracle):
    """Same facts as the reference's tests/test_pattern_utils.py: round trip and table size."""
    for gp, npat in (("NNMNN", 151875), ("SWSW", 81), ("NAA", 15)):
        n, nk, lvl = oracle.plan_info(gp)
        assert n == npat
        assert lvl == sum(len(oracle.CODE[c]) - 1 for c in gp)
        seen = set()
        for num in range(0, npat, max(1, npat // 500)):
            seen.add(oracle.num2pattern(gp, num))
        assert len(seen) == len(range(0, npat, max(1, npat // 500)))
        assert oracle.num2pattern(gp, npat - 1) == gp
        pn = oracle.kmer_patnums(gp)
        assert [oracle.num2pattern(gp, p) for p in pn[:50]] == oracle.kmers_of(gp)[:50]


@pytest.mark.parametrize("path", golden_files("single"), ids=lambda p: p.split("single_")[-1][:-4])
def test_single_dp_tables_match_reference(oracle, path):
    g = np.load(path)
    gp = str(g["gen_pat"])
    r = oracle.single_dp(gp, g["kmerM"], g["kmerU"], float(g["alpha"]), float(g["beta"]), float(g["penalty"]))
    assert np.array_equal(_bits(r["score"]), _bits(g["score"]))      # every float32 cell, bit-exact
    assert np.array_equal(r["M"], g["M"]) and np.array_equal(r["U"], g["U"])
    # backtrack pointer: reference stores the dense index of the c1 child, or self
    names = oracle.partition_names(gp, r["split"])
    assert names == [str(x) for x in g["names"]]
    assert np.array_equal((r["split"] == 0xFF), (g["bt"] == np.arange(len(g["bt"]), dtype=np.uint64)))
    assert _bits(r["score"][-1:])[0] == _bits(np.array([g["top_score"]], dtype=np.float32))[0]


@pytest.mark.parametrize("path", golden_files("cv"), ids=lambda p: p.split("cv_")[-1][:-4])
def test_cv_grid_matches_reference(oracle, path):
    g = np.load(path)
    gp, nf = str(g["gen_pat"]), int(g["nfolds"])
    alphas, pens = [float(a) for a in g["alphas"]], [float(c) for c in g["penalties"]]
    res = oracle.cv_grid(gp, g["kmerM"], g["kmerU"], alphas, pens, nf, int(g["seed"]))
    Mf, Uf = res["folds"]
    pn = oracle.kmer_patnums(gp)
    assert np.array_equal(g["M_folds"][pn], Mf) and np.array_equal(g["U_folds"][pn], Uf)   # RNG stream parity
    rows = "".join(f"{len(gp)} {a} {c} {str(t)}\n" for a, c, t in res["rows"])
    assert rows == str(g["cv_rows"])
    gi = 0
    for a_i in range(len(alphas)):
        for p_i in range(len(pens)):
            assert np.array_equal(_bits(g["train_tables"][gi][-1]), _bits(res["per_job"][(a_i, p_i)][0]))
            gi += 1
    assert np.array_equal(_bits(g["last_test_table"][-1]), _bits(res["per_job"][(len(alphas) - 1, len(pens) - 1)][1]))
    assert res["best"][0] == float(g["best_alpha"]) and res["best"][1] == float(g["best_penalty"])
    assert np.float32(res["best"][2]) == np.float32(g["best_test"])


def test_cv_full_table_one_fold(oracle):
    """Whole train table of one grid point, every fold, against the reference's score_mem."""
    path = [p for p in golden_files("cv") if "NNN_nb" in p][0]
    g = np.load(path)
    gp, nf = str(g["gen_pat"]), int(g["nfolds"])
    pn = oracle.kmer_patnums(gp)
    Mf, Uf = g["M_folds"][pn], g["U_folds"][pn]
    Mtot, Utot = Mf.sum(axis=1), Uf.sum(axis=1)
    Ms, Us = Mf.sum(axis=0), Uf.sum(axis=0)
    Mtr, Utr = Ms.sum() - Ms, Us.sum() - Us
    alphas, pens = list(g["alphas"]), list(g["penalties"])
    a_i, p_i = len(alphas) - 1, len(pens) - 1
    my = Mtr / (Mtr + Utr)
    betas = (alphas[a_i] * (1.0 - my)) / my
    for f in range(nf):
        train, test = oracle.cv_job(gp, Mtot, Utot, Mf[:, f], Uf[:, f], alphas[a_i], betas[f], pens[p_i])
        assert np.array_equal(_bits(train), _bits(g["train_tables"][-1][:, f]))
        assert np.array_equal(_bits(test), _bits(g["last_test_table"][:, f]))


def test_greedy_restatement_matches_reference_cli(oracle):
    """oracle.greedy (plain-Python restatement of greedy_res_kmer_table_ord) against the recorded reference run
    `--greedy -c 5 -a 0.8` on the 5-mer test data: same patterns in the same order, same float64 loss text."""
    import json
    import os

    from conftest import GOLDEN

    g = json.load(open(os.path.join(GOLDEN, "cli_5mers_greedy.json")))
    pos, bg = {}, {}
    for name, tab in (("mutated_5mers.txt", pos), ("background_5mers.txt", bg)):
        for line in open(os.path.join(GOLDEN, "data", name)):
            k, c = line.split()
            tab[k] = tab.get(k, 0) + int(c)
    gp = "NNMNN"
    km = oracle.kmers_of(gp)
    M = [pos.get(k, 0) for k in km]
    U = [bg.get(k, 0) - pos.get(k, 0) for k in km]
    alpha, pen = 0.8, 5.0
    mu = sum(M) / (sum(M) + sum(U))
    score, names, _, _ = oracle.greedy(gp, M, U, alpha, (alpha * (1.0 - mu)) / mu, pen)
    assert names == [line.split()[0] for line in g["stdout"].splitlines()[1:]]
    assert f"loss={score}" in g["stderr"]

