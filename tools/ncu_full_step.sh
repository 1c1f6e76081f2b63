#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): ncu --set full of the kernels of the THIRD full-graph training step
# (ML-25M shape, one GPU; 13 launches: 3 fwd layers, 4 bucket kernels, BPR user / neg / item passes, 3 bwd layers).
# The same program is run plain first.  TAG names the output files.
set -u
TAG=${1:-r2_full_step}
mkdir -p gpurun_out
python tools/prof_full_step.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
K='regex:rowtask_kernel|bpr_neg_kernel|bucket_|neg_hist'
ncu --set full --clock-control none --import-source on -k "$K" -s 26 -c 13 -f -o gpurun_out/${TAG} \
    python tools/prof_full_step.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --kernel-name regex:BprUserOwnOp > gpurun_out/${TAG}_source_user.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --kernel-name regex:FwdOp > gpurun_out/${TAG}_source_fwd.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/${TAG}.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/${TAG}.ncu-rep; echo "dropped ${TAG}.ncu-rep ($sz bytes)"; fi
ls -la gpurun_out | tail -8
