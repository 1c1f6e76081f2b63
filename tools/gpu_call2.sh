#!/bin/bash
# Runs ON THE GPU BOX (one GPU): round-2 second call -- the new tests, the bench line with the pipelined e2e, the
# C5 step on one GPU against float64, register-cap variants of the gather kernels, ncu of the scoring kernel.
set -u
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 $O/${TAG}_pytest_gpu.log)"
timeout 600 python bench.py --no-cpu > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
echo "bench rc=$? $(head -c 300 $O/${TAG}_bench_n1.json)"
python -c "import json;d=json.load(open('$O/${TAG}_bench_n1.json'));print('e2e',d['e2e']['ms_per_step'],d['e2e'].get('serial_ms_per_step'),'stages',d['stage_ms_per_step'])"
timeout 420 python bench.py --workload c5 --steps 3 --warmup 3 --check --no-cpu > $O/${TAG}_c5_bench_n1.json 2> $O/${TAG}_c5_bench_n1.err
echo "c5 rc=$? $(head -c 300 $O/${TAG}_c5_bench_n1.json)"
python -c "import json;d=json.load(open('$O/${TAG}_c5_bench_n1.json'));print('parity',d.get('parity'));print('stages',d['stage_ms_per_step']);print('roofline',d['roofline']['achieved'],d['roofline']['frac'])"
export LGCN_EDGE_CACHE=/dev/shm/lgcn_ab_edges.npy
V=movie-recommender-system-with-gnns_b200/csrc/build/variants
for lib in "" $V/liblgcn_bprminb3.so $V/liblgcn_bprminb4.so $V/liblgcn_spmmminb8.so $V/liblgcn_both.so; do
  LGCN_LIB_PATH=$lib timeout 200 python tools/time_sharded.py 2>&1 | grep -E "^world|^lib" >> $O/${TAG}_variants.txt
done
cat $O/${TAG}_variants.txt
python tools/prof_score.py tc > $O/${TAG}_score_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_topk_tc_kernel -s 3 -c 1 -f -o $O/${TAG}_score \
    python tools/prof_score.py tc > $O/${TAG}_score_ncu.log 2>&1
ncu -i $O/${TAG}_score.ncu-rep --page raw --csv > $O/${TAG}_score_raw.csv 2>/dev/null
grep -o "sm__pipe_tensor[^,]*" $O/${TAG}_score_raw.csv | head -3
for f in $O/${TAG}_*.err; do tail -c 20000 $f > $f.tail && mv $f.tail $f; done
ls -la $O | tail -15
