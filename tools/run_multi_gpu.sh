#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): the multi-GPU evidence of a round in one call.
#   bash tools/run_multi_gpu.sh N TAG [c5] [tests]
# bench line + full-size parity (--check: loss, final rows, dL/dE0 against float64) at N and N/2 GPUs; with `c5` the
# same on the 10x graph at N GPUs; with `tests` the small-shape sharded tests (CPU oracle) at world N first.
# Every command runs under its own timeout so that a hung rank cannot hold the box.
set -u
N=${1:-2}; TAG=${2:-r2}; shift 2
C5=0; TESTS=0
for a in "$@"; do [ "$a" = c5 ] && C5=1; [ "$a" = tests ] && TESTS=1; done
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29600
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
for n in $N $((N/2)); do
  [ $n -ge 2 ] || continue
  port=$((port+1))
  timeout 240 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --check > $O/${TAG}_bench_n${n}.json 2> $O/${TAG}_bench_n${n}.err
  echo "n=$n rc=$? $(head -c 330 $O/${TAG}_bench_n${n}.json)"
  python -c "import json;d=json.load(open('$O/${TAG}_bench_n${n}.json'));print('parity',d.get('parity'));print('stages',d.get('stage_ms_per_step'));print('e2e',d['e2e']['ms_per_step'],'prop',d['propagation']['ms'])"
  grep -E "NCCL INFO.*(nranks|NVLS)" $O/${TAG}_bench_n${n}.err | head -4 > $O/${TAG}_nccl_init_n${n}.txt
done
if [ $C5 = 1 ]; then
  port=$((port+1))
  timeout 420 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --workload c5 --steps 5 --check > $O/${TAG}_c5_bench_n${N}.json 2> $O/${TAG}_c5_bench_n${N}.err
  echo "c5 rc=$? $(head -c 400 $O/${TAG}_c5_bench_n${N}.json)"
  python -c "import json;d=json.load(open('$O/${TAG}_c5_bench_n${N}.json'));print('parity',d.get('parity'));print('stages',d.get('stage_ms_per_step'));print('roofline',d['roofline'])"
fi
if [ $TESTS = 1 ]; then
  timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -s -k "not world_one and not vs_fp64 and (${N}-fused or ${N}-nccl or scoring and ${N})" 2>&1 | grep -v "^$" | tail -30 > $O/${TAG}_sharded_tests_n${N}.log
  tail -4 $O/${TAG}_sharded_tests_n${N}.log
fi
for f in $O/${TAG}_*.err; do tail -c 20000 $f > $f.tail && mv $f.tail $f; done
