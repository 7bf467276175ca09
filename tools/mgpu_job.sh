# 8-GPU evidence run: bench at N = 8 and N = 4, sharded single DP with one- and two-dimensional ownership
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $R --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_b_n8.json 2> gpurun_out/r2_b_n8.err; echo "bench8 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 200 $R --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_b_n4.json 2> gpurun_out/r2_b_n4.err; echo "bench4 rc=$?"
for n in 8 4; do for m in 0 1; do
  if [ $m = 1 ]; then export KP_SHARD_1D=1; else unset KP_SHARD_1D; fi
  timeout 120 $R --nproc-per-node $n --master-port 2955$n tests/mgpu_sharded_check.py NNNNANNNN 6 1 2>/dev/null | tail -2 | sed "s/^/n=$n 1D=$m: /"
done; done | tee gpurun_out/r2_shard_ab.txt
unset KP_SHARD_1D
python -c "
import json
for n in (8,4):
    d=json.load(open('gpurun_out/r2_b_n%d.json'%n)); print(n, d['ms_per_step'], d['parity_checked'], d['sharded_single_dp']['ms'], d['sharded_single_dp']['parity_checked'])"
