#!/bin/bash
# Runs ON THE GPU BOX (one GPU): per-kernel times of the full-graph step (ncu launch list of tools/prof_full_step.py,
# plain run first).   bash tools/gpu_launchlist.sh TAG
set -u
TAG=${1:-ll}
O=gpurun_out
mkdir -p $O
LGCN_PROF_STEPS=4 python tools/prof_full_step.py > $O/${TAG}_plain.log 2>&1 || { tail -5 $O/${TAG}_plain.log; exit 1; }
LGCN_PROF_STEPS=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv \
    python tools/prof_full_step.py > $O/${TAG}_ncu.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv 24 > $O/${TAG}_launch_summary.txt
cat $O/${TAG}_launch_summary.txt
