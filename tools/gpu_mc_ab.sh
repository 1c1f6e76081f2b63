#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): stage times of the sharded step with NVLS multicast stores vs per-peer stores.
set -u
N=${1:-8}; TAG=${2:-mc}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29700
for mc in 1 0; do
  port=$((port+1))
  echo "== LGCN_P2P_MULTICAST=$mc" >> $O/${TAG}_time_sharded_n${N}.txt
  LGCN_P2P_MULTICAST=$mc timeout 200 $TR --nproc-per-node $N --master-port $port tools/time_sharded.py 2>&1 | grep -E "^world|^shard|^lib|Error" >> $O/${TAG}_time_sharded_n${N}.txt
done
cat $O/${TAG}_time_sharded_n${N}.txt
