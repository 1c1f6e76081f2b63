#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): ncu captures for profiles/.  The same command is run plain first.
# Reports are exported to CSV here because .ncu-rep files can exceed the 64 MiB copy-back limit.
set -u
mkdir -p gpurun_out
K='regex:rowtask_kernel|clip_adam_kernel|inactive_kernel'
python tools/prof_step.py > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
# (a) second full-graph training step: 9 launches (3 fwd, BPR A/B, 3 bwd, clip_adam)
ncu --set full --clock-control none --import-source on -k "$K" -s 9 -c 9 -f -o gpurun_out/full_step \
    python tools/prof_step.py > gpurun_out/prof_ncu_a.log 2>&1
# (b) second median Cluster-GCN batch step: 11 launches (+ the two inactive-row kernels)
ncu --set full --clock-control none -k "$K" -s 51 -c 11 -f -o gpurun_out/cluster_step \
    python tools/prof_step.py > gpurun_out/prof_ncu_b.log 2>&1
for r in full_step cluster_step; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/${r}_details.csv 2>/dev/null
done
ncu -i gpurun_out/full_step.ncu-rep --page source --csv --kernel-name regex:rowtask_kernel > gpurun_out/full_step_source.csv 2>/dev/null
ls -la gpurun_out
du -sm gpurun_out
# keep the copy-back under the limit
for r in full_step cluster_step; do
  sz=$(stat -c %s gpurun_out/$r.ncu-rep); if [ "$sz" -gt 25000000 ]; then rm gpurun_out/$r.ncu-rep; echo "dropped $r.ncu-rep ($sz bytes)"; fi
done
