# NVLink bytes per rank of the sharded single DP under the three ownership schemes (8 GPUs), and the default at 4 and 2
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for cfg in "8 tables" "8 modular" "8 1d" "4 tables" "2 tables"; do set -- $cfg
  unset KP_SHARD_1D KP_SHARD_MODULAR
  [ $2 = modular ] && export KP_SHARD_MODULAR=1
  [ $2 = 1d ] && export KP_SHARD_1D=1
  timeout 120 $R --nproc-per-node $1 --master-port 2957$1 tests/mgpu_sharded_check.py NNNNANNNN 6 1 2>&1 | grep -E "^NNNN|SHARDED|NVLink" | sed -E "s/^/n=$1 $2: /; s/NNNNANNNN: npat 2562890625, //; s/, partition.*//"
done | tee gpurun_out/r02_nvlink_counters.txt
