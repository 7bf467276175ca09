"""CPU tests of the host side: IUPAC tables, input parsing, fold sampler, CV reduction, job sharding,
the C ABI surface.  No GPU needed."""
import ctypes
import io
import json
import os
import re
import types

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_files


def test_iupac_tables_against_oracle(oracle):
    from kmerpapa_b200 import iupac

    for gp in ("NNMNN", "SWSW", "RYNBA", "NAA"):
        npat, nk, lvl = oracle.plan_info(gp)
        PE = iupac.PatternEnumeration(gp)
        assert PE.npat == npat == iupac.pattern_max(gp)
        assert iupac.pattern_level(gp) == lvl
        assert iupac.matches(gp) == oracle.kmers_of(gp) and len(iupac.matches(gp)) == nk
        for num in list(range(0, npat, max(1, npat // 300))) + [npat - 1]:
            pat = oracle.num2pattern(gp, num)
            assert PE.num2pattern(num) == pat and PE.pattern2num(pat) == num
        assert iupac.lca_pattern(iupac.matches(gp)) == gp
    assert iupac.contains("NNANN", "CGATT") and not iupac.contains("NNANN", "CGCTT")
    assert iupac.kmer_code("ACGT") == 1 | 2 << 4 | 4 << 8 | 8 << 12


def test_read_input_test_data():
    from kmerpapa_b200 import io_utils

    args = types.SimpleNamespace(positive=open(f"{GOLDEN}/data/mutated_5mers.txt"), negative=None,
                                 background=open(f"{GOLDEN}/data/background_5mers.txt"), joint_context_counts=None)
    table, n_unmut, n_mut = io_utils.read_input(args, None)
    assert (n_mut, n_unmut) == (59479, 2164774234)           # the reference's "Input data read." line (cfg 1)
    assert len(table) == 512 and all(u >= 0 for _, u in table.values())
    args = types.SimpleNamespace(positive=open(f"{GOLDEN}/data/mutated_5mers.txt"), negative=None,
                                 background=open(f"{GOLDEN}/data/background_7mers.txt"), joint_context_counts=None)
    t2, u2, m2 = io_utils.read_input(args, "RNAYN")           # longer background k-mers are centre-trimmed
    assert all(len(k) == 5 and k[0] in "AG" and k[2] == "A" and k[3] in "CT" for k in t2)


def test_read_dict_rules():
    from kmerpapa_b200 import io_utils

    f = io.StringIO("ACGTA 3\nACNTA 9\nACGTA 2.0\nTTTTT 1e1\n")
    table, total = io_utils.read_dict(f, None)
    assert table == {"ACGTA": 5, "TTTTT": 10} and total == 15     # N skipped, duplicates add, float counts
    f = io.StringIO("AACGTAA 4\nCACGTAC 1\n")
    table, total = io_utils.read_dict(f, None, length=5)
    assert table == {"ACGTA": 5} and total == 5
    with pytest.raises(AssertionError):
        io_utils.read_dict(io.StringIO("ACG -1\n"), None)
    joint, n_unmut, n_mut = io_utils.read_joint_kmer_counts(io.StringIO("ACG 2 10\nACT 0 5\n"), None)
    assert joint == {"ACG": (2, 8), "ACT": (0, 5)} and (n_unmut, n_mut) == (13, 2)
    small, gp = io_utils.downsize_contextD({"AACGT": (1, 2), "CACGA": (3, 4)}, "NNCGN", 3)
    assert small == {"ACG": [4, 6]} and gp == "NCG"


@pytest.mark.parametrize("path", golden_files("cv"), ids=lambda p: p.split("cv_")[-1][:-4])
def test_fold_sampler_matches_reference_stream(oracle, path):
    """Same numpy RandomState stream as the reference's CV_tools (held-out counts per fold)."""
    from kmerpapa_b200 import CV_tools, iupac

    g = np.load(path)
    gp, nf = str(g["gen_pat"]), int(g["nfolds"])
    kmers = iupac.matches(gp)
    Mf, Uf = CV_tools.sample_fold_counts(kmers, g["kmerM"], g["kmerU"], nf, np.random.RandomState(int(g["seed"])))
    pn = oracle.kmer_patnums(gp).astype(np.int64)
    assert np.array_equal(Mf, g["M_folds"][pn]) and np.array_equal(Uf, g["U_folds"][pn])
    # a shuffled input order must give the same table (the sampler sorts k-mers itself)
    perm = np.random.default_rng(0).permutation(len(kmers))
    Mf2, Uf2 = CV_tools.sample_fold_counts([kmers[i] for i in perm], g["kmerM"][perm], g["kmerU"][perm], nf,
                                           np.random.RandomState(int(g["seed"])))
    assert np.array_equal(Mf2, Mf[perm]) and np.array_equal(Uf2, Uf[perm])


def test_streaming_sampler_equals_batch_sampler():
    """iter_fold_counts (one fold at a time, feeds the GPU while the next fold is drawn) == sample_fold_counts."""
    from kmerpapa_b200 import CV_tools, iupac

    kmers = iupac.matches("NNRN")
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(kmers))
    kmers = [kmers[i] for i in perm]
    pos, neg = rng.integers(0, 50, len(kmers)), rng.integers(0, 5000, len(kmers))
    for nf in (2, 3, 5):
        Mf, Uf = CV_tools.sample_fold_counts(kmers, pos, neg, nf, np.random.RandomState(5))
        seen = []
        for f, M, U in CV_tools.iter_fold_counts(kmers, pos, neg, nf, np.random.RandomState(5)):
            seen.append(f)
            assert np.array_equal(M, Mf[:, f]) and np.array_equal(U, Uf[:, f])
        assert seen == list(range(nf))


def test_pipelined_grid_equals_batch_grid():
    """run_grid with a runner that accepts folds one at a time (the GPU runner's interface) gives the results of
    the batch path, H7 quirk of later iterations included."""
    from kmerpapa_b200 import iupac
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv

    class Batch:
        def set_folds(self, Mf, Uf):
            self.Mf, self.Uf = Mf, Uf

        def run(self, f, alpha, beta, penalty):   # a stand-in "DP": any deterministic function of its inputs
            x = float(self.Mf[:, f].sum()) * alpha + float(self.Uf[:, f].sum()) * 1e-3 + beta * penalty
            return np.float32(x), np.float32(x / 3)

    class Streaming(Batch):
        def begin_folds(self, nkmer, nfolds):
            self.Mf = np.zeros((nkmer, nfolds), dtype=np.uint64)
            self.Uf = np.zeros((nkmer, nfolds), dtype=np.uint64)

        def set_fold(self, f, M, U):
            self.Mf[:, f], self.Uf[:, f] = M, U

    gp = "NNMN"
    kmers = iupac.matches(gp)
    rng = np.random.default_rng(1)
    pos, neg = rng.integers(0, 30, len(kmers)), rng.integers(100, 9000, len(kmers))
    codes = iupac.kmer_codes(kmers)
    for nit in (1, 3):
        a = cv.run_grid(gp, kmers, codes, pos, neg, [0.5, 2.0], [3.0, 5.0], 4, nit, 9, runner=Batch(), gather_device=None)
        b = cv.run_grid(gp, kmers, codes, pos, neg, [0.5, 2.0], [3.0, 5.0], 4, nit, 9, runner=Streaming(), gather_device=None)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_fold_sums_like_reference_test():
    """tests/test_CV_tools.py of the reference: folds add back to the table."""
    from kmerpapa_b200 import CV_tools

    kmers = ["AAA", "CAA", "GAA", "TAA"]
    pos, neg = np.array([10, 200, 500, 300]), np.array([100, 1000, 2000, 1000])
    Mf, Uf = CV_tools.sample_fold_counts(kmers, pos, neg, 10, np.random.RandomState(0))
    assert Mf.shape == (4, 10) and np.array_equal(Mf.sum(axis=1), pos) and np.array_equal(Uf.sum(axis=1), neg)


def test_cv_reduction_and_selection_format():
    """float32 sequential fold sum, strict '<' selection alpha-outer / penalty-inner, CVfile row text."""
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv

    res = np.zeros((1, 2, 2, 2, 2), dtype=np.float32)
    res[0, :, 0, 0, 1] = [662892.56, 662892.56]
    res[0, :, 0, 1, 1] = [662838.5, 662838.5]
    res[0, :, 1, 0, 1] = [662838.5, 662838.5]       # tie with the earlier grid point: the earlier one stays
    res[0, :, 1, 1, 1] = [662900.0, 662900.0]
    out = io.StringIO()
    a, c, best = cv.select_best([0.8, 1.0], [3.0, 5.0], res, 1, 2, 5, out)
    assert (a, c) == (0.8, 5.0) and isinstance(best, np.float32) and best == np.float32(1325677.0)
    assert out.getvalue().splitlines()[1] == "5 0.8 5.0 1.325677e+06"
    assert cv.fold_sum([np.float32(16777216.0), np.float32(1.0), np.float32(1.0)], 1) == np.float32(16777216.0)  # f32 accumulation


def test_job_sharding():
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv

    jobs = cv.job_list(1, 5, 3, 3)
    assert len(jobs) == 45 and jobs[0] == (0, 0, 0, 0) and jobs[9] == (0, 1, 0, 0)
    for world in (1, 2, 4, 8, 45, 64):
        covered = []
        for r in range(world):
            lo, hi = cv.shard_bounds(45, r, world)
            covered += list(range(lo, hi))
            folds = {jobs[j][1] for j in range(lo, hi)}
            if world <= 8:
                assert len(folds) <= 2 + (45 // world) // 9
        assert covered == list(range(45))
    assert cv.covering_patterns_per_kmer("NNMNA") == 8 * 8 * 2 * 8


def test_cli_smoke(capsys):
    """The reference's own CLI tests: no input -> help + 'input error', exit code 0; -h exits."""
    from kmerpapa_b200 import cli

    assert cli.main([]) == 0
    with pytest.raises(SystemExit):
        cli.main(["-h"])
    assert "kmerpapa" in capsys.readouterr().out
    p = cli.get_parser().parse_args(["--n_folds", "5", "-c", "3", "5"])
    assert p.nfolds == 5 and p.penalty_values == [3.0, 5.0] and p.pseudo_counts == [0.8]
    assert cli.get_parser().parse_args(["--nfolds", "3"]).nfolds == 3


def test_c_abi_exports_every_declared_symbol():
    """libkpapa.so loads and exports exactly what include/kmerpapa_b200.h declares (no compute calls)."""
    from kmerpapa_b200 import _native, build

    build.build_native()
    header = open(os.path.join(ROOT, "include", "kmerpapa_b200.h")).read()
    declared = set(re.findall(r"\b(kp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.SYMBOLS), "ctypes table and header disagree"
    assert _native.lib().kp_version() >= 100
    assert ctypes.sizeof(_native.PlanInfo) == 7 * 8 + 12 * 4


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path raises; it never routes to the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kmerpapa_b200._native import KpError
    from kmerpapa_b200.engine import get_plan

    with pytest.raises(KpError):
        get_plan("NNN")
    src = "".join(open(os.path.join(dp, f)).read() for dp, _, fs in os.walk(os.path.join(ROOT, "kmerpapa_b200"))
                  for f in fs if f.endswith(".py"))
    assert "oracle" not in src.replace("no CPU", ""), "product code must not reference the oracle"


def test_sublattice_helpers_against_the_string_enumeration():
    from fullsize_util import sub_kmer_select, sublattice_patnums
    from kmerpapa_b200 import iupac

    gen, sub = "NRANY", "SAAKY"
    PE, PS = iupac.PatternEnumeration(gen), iupac.PatternEnumeration(sub)
    want = np.array([PE.pattern2num(PS.num2pattern(i)) for i in range(PS.npat)], dtype=np.uint64)
    assert np.array_equal(sublattice_patnums(gen, sub), want)
    kf = iupac.matches(gen)
    assert [kf[i] for i in sub_kmer_select(gen, sub)] == iupac.matches(sub)


def test_a_sublattice_is_a_dp_of_its_own(oracle):
    """The premise of the full-size parity tests, checked on the oracle alone: the table of the sub-patterns of S
    (scores and split decisions) sits verbatim inside the table of any general pattern that contains S."""
    from fullsize_util import sub_kmer_select, sublattice_patnums

    gen = "NNANN"
    rng = np.random.default_rng(11)
    U = 1 + rng.negative_binomial(2, 2 / (2 + 800.0), size=256)
    M = rng.binomial(U, 0.05)
    full = oracle.single_dp(gen, M, U, 0.7, 30.0, 3.0)
    for sub in ("ANANN", "NNANT", "SKACN"):
        sel = sub_kmer_select(gen, sub)
        ref = oracle.single_dp(sub, M[sel], U[sel], 0.7, 30.0, 3.0)
        nums = sublattice_patnums(gen, sub).astype(np.int64)
        assert np.array_equal(full["score"][nums].view(np.uint32), ref["score"].view(np.uint32))
        assert np.array_equal(full["split"][nums], ref["split"])
        assert np.array_equal(full["M"][nums], ref["M"])


def test_bench_config_is_shared_by_both_arms_and_goldens_are_checked():
    """bench.py: both arms print the same `config` object; a result that differs from the full-size golden aborts."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("kp_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in ("cfg3", "cfg4", "cfg5", "n9m"):
        a, b = bench.config_for(name, 1), bench.config_for(name, 1)
        assert a == b and json.loads(json.dumps(a)) == a and "workload" in a and "model" not in a
    assert bench.config_for("cfg3", 1)["npat"] == 2562890625 and bench.config_for("cfg5", 1)["npat"] == 922640625
    assert bench.host_threads() >= 1
    g = bench.golden("cfg4")
    assert g is not None and len(g["jobs_run"]) == 45 and g["selected"][:2] == [10.0, 6.0]
    res = np.array([int(x, 16) for x in g["job_bits"]], dtype=np.uint32).view(np.float32).reshape(1, 5, 3, 3, 2)
    best = (g["selected"][0], g["selected"][1], np.float32(g["selected"][2]))
    out = bench.check_cv_against_golden(res, best)
    assert out["checked"] and out["jobs_compared"] == 45 and out["selected_compared"]
    bad = res.copy()
    bad[0, 2, 1, 1, 1] = np.nextafter(bad[0, 2, 1, 1, 1], np.float32(0))
    with pytest.raises(AssertionError):
        bench.check_cv_against_golden(bad, best)
    g3 = bench.golden("cfg3")
    with pytest.raises(AssertionError):
        bench.check_single_against_golden("cfg3", np.float32(g3["loss"]) + np.float32(8.0), np.zeros(g3["partition_patterns"], dtype=np.uint64))
    assert bench.check_single_against_golden("n9m", np.float32(1.0), np.zeros(1, dtype=np.uint64))["checked"] is False
