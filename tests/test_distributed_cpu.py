"""Job sharding + gather over torch.distributed with the gloo backend, world_size 2, on CPU.
The per-job runner is a stand-in (the CPU oracle on a small general pattern): what is under test is the
host logic every rank runs — the fold sampler one fold ahead of the jobs (the runner has the GPU runner's
streaming interface), contiguous job chunks, one all_gather, identical selection."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT

GEN_PAT, ALPHAS, PENS, NF, SEED = "NNN", [0.5, 2.0], [2.0, 5.0, 7.0], 3, 4


def _data():
    rng = np.random.default_rng(3)
    U = 1 + rng.negative_binomial(2, 2 / (2 + 900.0), size=64)
    M = rng.binomial(U, 0.03)
    return M.astype(np.int64), U.astype(np.int64)


class OracleRunner:
    device = None

    def __init__(self, gen_pat, M, U):
        sys.path.insert(0, ROOT)
        from oracle import kp_oracle

        self.O, self.gp, self.M, self.U = kp_oracle, gen_pat, M, U
        self.calls = 0

    def set_folds(self, Mf, Uf):
        self.Mf, self.Uf = Mf, Uf

    # the GPU runner's streaming interface: folds arrive one at a time from the sampler thread
    def begin_folds(self, nkmer, nfolds):
        self.Mf = np.zeros((nkmer, nfolds), dtype=np.uint64)
        self.Uf = np.zeros((nkmer, nfolds), dtype=np.uint64)

    def set_fold(self, f, M, U):
        self.Mf[:, f], self.Uf[:, f] = M, U

    def run(self, f, alpha, beta, penalty):
        self.calls += 1
        tr, te = self.O.cv_job(self.gp, self.M, self.U, self.Mf[:, f], self.Uf[:, f], alpha, beta, penalty, nthreads=1)
        return tr[-1], te[-1]


def _grid(rank, world):
    from kmerpapa_b200 import iupac
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv

    M, U = _data()
    kmers = iupac.matches(GEN_PAT)
    runner = OracleRunner(GEN_PAT, M, U)
    res = cv.run_grid(GEN_PAT, kmers, None, M, U, ALPHAS, PENS, NF, 1, SEED, runner=runner, gather_device=None)
    return res, runner.calls


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res, calls = _grid(rank, world)
    q.put((rank, res.tobytes(), calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_cv_grid_sharded_over_two_ranks_matches_single_process(oracle):
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv

    single, calls = _grid(0, 1)
    njobs = NF * len(ALPHAS) * len(PENS)
    assert calls == njobs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, blob, c in got:
        assert blob == single.tobytes()                      # every rank ends with the full, identical result
        lo, hi = cv.shard_bounds(njobs, rank, 2)
        assert c == hi - lo                                  # and only ran its own chunk
    # selection equals the oracle's own grid driver
    M, U = _data()
    ref = oracle.cv_grid(GEN_PAT, M, U, ALPHAS, PENS, NF, SEED, nthreads=1)
    a, c, best = cv.select_best(ALPHAS, PENS, single, 1, NF, len(GEN_PAT))
    assert (a, c) == (ref["best"][0], ref["best"][1]) and np.float32(best) == np.float32(ref["best"][2])


def _seed_worker(rank, world, port, q, seed):
    import argparse

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kmerpapa_b200 import cli

    args = argparse.Namespace(seed=seed if rank == 0 else (None if seed is None else seed + 17))
    cli.share_seed(args)
    q.put((rank, args.seed))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("seed", [None, 5])
def test_ranks_share_one_seed(seed):
    """Without --seed every rank would seed its fold sampler from OS entropy and cross-validate on different folds:
    rank 0's seed (drawn when unset) is broadcast."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + (0 if seed is None else 1)
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, q, seed)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1] and isinstance(got[0], int)
    if seed is not None:
        assert got[0] == seed
    np.random.RandomState(got[0])   # a valid legacy seed


class PipelinedOracleRunner(OracleRunner):
    """The GPU runner's pipelined interface (submit / flush: results come back one or two jobs late) on top of the oracle."""

    DEPTH = 2

    def __init__(self, *a):
        super().__init__(*a)
        self.pending = []

    def submit(self, tag, f, alpha, beta, penalty):
        self.pending.append((tag, self.run(f, alpha, beta, penalty)))
        out = []
        while len(self.pending) >= self.DEPTH:
            tag0, (tr, te) = self.pending.pop(0)
            out.append((tag0, tr, te))
        return out

    def flush(self):
        out = [(tag, tr, te) for tag, (tr, te) in self.pending]
        self.pending = []
        return out


@pytest.mark.timeout(300)
@pytest.mark.parametrize("presample", [False, True])
def test_pipelined_runner_fills_the_same_result_slots(presample):
    """run_grid with a runner whose results arrive late (the GPU runner queues the next DP before it reads the previous
    job's losses) must file every result under its own job, on the streaming-sampler path and on the presampled one."""
    from kmerpapa_b200 import CV_tools, iupac
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv

    M, U = _data()
    kmers = iupac.matches(GEN_PAT)
    pre = None
    if presample:
        pre = [CV_tools.sample_fold_counts(kmers, M, U, NF, np.random.RandomState(SEED))]
    out = []
    for cls in (OracleRunner, PipelinedOracleRunner):
        runner = cls(GEN_PAT, M, U)
        out.append(cv.run_grid(GEN_PAT, kmers, None, M, U, ALPHAS, PENS, NF, 1, SEED, runner=runner, gather_device=None,
                               presampled=pre))
        assert runner.calls == NF * len(ALPHAS) * len(PENS)
    assert out[0].tobytes() == out[1].tobytes()
