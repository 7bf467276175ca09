"""Run under torchrun on N GPUs: one DP sharded over the ranks (CUDA IPC peer pointers, NCCL barrier between waves)
must give the partition, the top score and the split codes of the unsharded DP computed on rank 0.
Usage: torchrun --nproc-per-node N tests/mgpu_sharded_check.py [gen_pat] [reps] [replicate 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from kmerpapa_b200 import sharded, synthetic
from kmerpapa_b200.engine import get_plan


def nvlink_bytes(index):
    """(received, sent) bytes over all NVLink links of physical GPU `index` so far (NVML field values), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        rx = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])[0]
        tx = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF)])[0]
        if rx.nvmlReturn != 0 or tx.nvmlReturn != 0:
            return None
        return int(rx.value.ullVal) * 1024, int(tx.value.ullVal) * 1024     # the counters are in KiB
    except Exception:
        return None


def main():
    gen_pat = sys.argv[1] if len(sys.argv) > 1 else "NNNNANNN"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    replicate = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kmers, pos, neg = synthetic.negbin_counts(gen_pat, 4242)
    plan = get_plan(gen_pat, local)
    kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
    eM, eU = plan.expand(kM, kU)
    mc = int(pos.sum() + neg.sum())
    mu = int(pos.sum()) / mc
    alpha, penalty = 1.0, 6.0
    beta = alpha * (1 - mu) / mu
    sh = sharded.ShardedDP(plan, rank, world, replicate=replicate)
    sh.connect()
    ms = []
    nv0 = nvlink_bytes(local)
    for rep in range(reps):
        sh.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sh.run(eM, eU, mc, alpha, beta, penalty)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=plan.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    nv1 = nvlink_bytes(local)
    if nv0 is not None and nv1 is not None:   # NVLink traffic of the DPs alone (the reads of the backtrack come after)
        t = torch.tensor([(nv1[0] - nv0[0]) / reps, (nv1[1] - nv0[1]) / reps], dtype=torch.float64, device=plan.device)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        if rank == 0:
            rx = [float(x[0]) / 1e9 for x in allt]
            tx = [float(x[1]) / 1e9 for x in allt]
            print("NVLink GB per DP and rank (NVML throughput counters): received " + " ".join(f"{x:.2f}" for x in rx) +
                  " | sent " + " ".join(f"{x:.2f}" for x in tx) + f" | busiest receiver {max(rx):.2f} GB", flush=True)
    part = sh.backtrack()
    top = sh.top_score()
    # size-independent properties of the result (the only checks available when the table exceeds one GPU):
    # the partition covers every k-mer exactly once, its counts add up to the totals, and the loss of the
    # general pattern is the sum of the leaves' self-scores (float32 tree sum vs float64 here: 1e-5 relative)
    from kmerpapa_b200 import iupac

    PE = iupac.PatternEnumeration(gen_pat)
    names = [PE.num2pattern(int(x)) for x in part]
    nk = sum(int(np.prod([len(iupac.matches(c)) for c in nm])) for nm in names)
    M, U = plan.pattern_counts(kM, kU, part)
    prop = (nk == len(kmers) and int(M.sum()) == int(pos.sum()) and int(U.sum()) == int(neg.sum()))
    Md, Ud = M.astype(np.float64), U.astype(np.float64)
    pr = (Md + alpha) / (Md + Ud + alpha + beta)
    with np.errstate(divide="ignore", invalid="ignore"):
        leaf = penalty - 2.0 * (np.where(M > 0, Md * np.log(pr), 0.0) + np.where(U > 0, Ud * np.log1p(-pr), 0.0))
    prop = prop and abs(float(leaf.sum()) - float(top)) <= 1e-5 * abs(float(top))
    rng = np.random.default_rng(11)
    pats = np.unique(np.concatenate([rng.integers(0, plan.npat, size=50000, dtype=np.uint64), part]))
    vals, flags, codes = sh.gather(pats, codes=True)
    ok = bool(prop)
    if rank == 0:
        print(f"properties: k-mers covered {nk}/{len(kmers)}, counts {int(M.sum())}/{int(pos.sum())} {int(U.sum())}/{int(neg.sum())}, "
              f"sum of leaf scores {leaf.sum():.1f} vs top {float(top):.1f} -> {'ok' if prop else 'FAIL'}", flush=True)
        fits = plan.info.table_elems * 4 < 150e9
        if fits:
            best, kept = plan.dp_single(eM, eU, mc, alpha, beta, penalty)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            best, kept = plan.dp_single(eM, eU, mc, alpha, beta, penalty)
            e1.record()
            torch.cuda.synchronize()
            single_ms = e0.elapsed_time(e1)
            ref_part = plan.backtrack(best, kept)
            ref_codes = plan.split_codes(best, kept, pats)
            ok = ok and (np.array_equal(part, ref_part) and plan.top_score(best) == top and np.array_equal(codes, ref_codes))
            print(f"{gen_pat}: npat {plan.npat}, world {world}, {'replicated' if replicate else 'partitioned'}: sharded {min(ms):.3f} ms ({plan.npat / min(ms) / 1e6:.1f} Gpat/s), "
                  f"one GPU {single_ms:.3f} ms, speed-up {single_ms / min(ms):.2f}x, partition {len(part)} patterns, "
                  f"top {top}", flush=True)
        else:
            print(f"{gen_pat}: npat {plan.npat}, world {world}: sharded {min(ms):.3f} ms ({plan.npat / min(ms) / 1e6:.1f} Gpat/s), "
                  f"partition {len(part)} patterns, top {top} (too large for one GPU: no unsharded comparison)", flush=True)
    # every rank must see the same results through its own peer mappings
    blob = [None] * world
    dist.all_gather_object(blob, (float(top), part.tobytes(), codes.tobytes()))
    ok = ok and all(b == blob[0] for b in blob)
    flag = torch.tensor([1 if ok else 0], device=plan.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED OK" if int(flag.item()) == 1 else "SHARDED MISMATCH", flush=True)
    sh.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
