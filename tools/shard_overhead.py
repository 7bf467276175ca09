import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
from kmerpapa_b200 import sharded, synthetic
from kmerpapa_b200.engine import get_plan
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gp = sys.argv[1] if len(sys.argv) > 1 else "NNNNANNNN"
kmers, pos, neg = synthetic.negbin_counts(gp, 9003)
plan = get_plan(gp, local)
kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
eM, eU = plan.expand(kM, kU)
mc = int(pos.sum() + neg.sum()); mu = int(pos.sum()) / mc
for rep in range(3):
    for replicate in (True, False):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        sh = sharded.ShardedDP(plan, rank, world, replicate=replicate)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        sh.connect()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        sh.run(eM, eU, mc, 1.0, (1 - mu) / mu, 6.0)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        top = sh.top_score(); part = sh.backtrack()
        t4 = time.perf_counter()
        sh.close()
        torch.cuda.synchronize(); t5 = time.perf_counter()
        if rank == 0:
            print(f"{gp} replicate={replicate}: create {1e3*(t1-t0):.1f} connect {1e3*(t2-t1):.1f} run {1e3*(t3-t2):.1f} "
                  f"read {1e3*(t4-t3):.1f} close {1e3*(t5-t4):.1f} ms", flush=True)
dist.destroy_process_group()
