#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): the multi-GPU evidence of a round in one call.
#   bash tools/run_multi_gpu.sh N TAG [c5] [quick]
# sharded tests at world N (small shapes vs the CPU oracle, ML-25M vs float64), stage timing and bench line at N GPUs
# (and N/2); with `c5` also the 10x graph: bench line + parity against float64 in one process group.
set -u
N=${1:-2}; TAG=${2:-r2}; C5=${3:-}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_sharded.py -x -q -s -k "not world_one and (${N}-fused or ${N}-nccl or $((N/2))-fused or vs_fp64 and ${N})" 2>&1 | grep -v "^$" | tail -30 > $O/${TAG}_sharded_tests_n${N}.log
tail -4 $O/${TAG}_sharded_tests_n${N}.log
port=29600
for n in $N $((N/2)); do
  [ $n -ge 2 ] || continue
  port=$((port+1))
  $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --check > $O/${TAG}_bench_n${n}.json 2> $O/${TAG}_bench_n${n}.err
  echo "n=$n rc=$? $(head -c 330 $O/${TAG}_bench_n${n}.json)"
done
port=$((port+1))
$TR --nproc-per-node $N --master-port $port tools/time_sharded.py 2>&1 | grep -E "^world|^shard" > $O/${TAG}_time_sharded_n${N}.txt
cat $O/${TAG}_time_sharded_n${N}.txt
if [ "$C5" = "c5" ]; then
  port=$((port+1))
  $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --workload c5 --steps 5 --check > $O/${TAG}_c5_bench_n${N}.json 2> $O/${TAG}_c5_bench_n${N}.err
  echo "c5 rc=$? $(head -c 400 $O/${TAG}_c5_bench_n${N}.json)"
  python -c "import json;d=json.load(open('$O/${TAG}_c5_bench_n${N}.json'));print(d.get('parity'));print(d.get('stage_ms_per_step'))"
fi
