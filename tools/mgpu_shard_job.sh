# sharded single DP on 8 and 4 GPUs: optimised two-dimensional ownership (default), modular two-dimensional, one-dimensional
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do for m in opt modular 1d; do
  unset KP_SHARD_1D KP_SHARD_MODULAR
  [ $m = modular ] && export KP_SHARD_MODULAR=1
  [ $m = 1d ] && export KP_SHARD_1D=1
  KP_SHARD_VERBOSE=1 timeout 120 $R --nproc-per-node $n --master-port 2955$n tests/mgpu_sharded_check.py NNNNANNNN 6 1 2>&1 | grep -E "^NNNN|SHARDED|busiest" | sort -u | sed "s/^/n=$n $m: /"
done; done | tee gpurun_out/r2_shard_ab2.txt
