#!/bin/bash
# Runs ON THE GPU BOX (gpurun, one GPU): the single-GPU evidence of a round in one call.
#   bash tools/gpu_round_check.sh TAG [notests]
# pytest -m gpu, the default bench line, the ncu launch list of the same bench command, ncu --set full of one
# full-graph training step.  Every ncu pass runs only after the plain command exited 0.
set -u
TAG=${1:-r2}; NOTESTS=${2:-}
O=gpurun_out
mkdir -p $O
if [ "$NOTESTS" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -q --durations=15 > $O/${TAG}_pytest_gpu.log 2>&1
  echo "pytest rc=$? $(tail -1 $O/${TAG}_pytest_gpu.log)"
fi
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
rc=$?; echo "bench rc=$rc $(head -c 400 $O/${TAG}_bench_n1.json)"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu > $O/${TAG}_launches_ncu.log 2>&1
  python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launch_summary.txt 2>&1
  head -30 $O/${TAG}_launch_summary.txt
fi
timeout 600 bash tools/ncu_full_step.sh ${TAG}_full_step
