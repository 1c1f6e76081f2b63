#!/bin/bash
# Runs ON THE GPU BOX (one GPU): A/B of library variants (lgcn_b200.build.build_variant) on the full-graph step.
#   bash tools/gpu_variants.sh TAG variant_tag ...     ("" = the default library)
set -u
TAG=${1:-ab}; shift
O=gpurun_out
mkdir -p $O
export LGCN_EDGE_CACHE=/dev/shm/lgcn_ab_edges.npy
V=movie-recommender-system-with-gnns_b200/csrc/build/variants
for tag in default "$@"; do
  lib=""; [ "$tag" != default ] && lib=$V/liblgcn_$tag.so
  LGCN_LIB_PATH=$lib timeout 200 python tools/time_sharded.py 2>&1 | grep -E "^world|^lib|Error|error" >> $O/${TAG}_variants.txt
done
cat $O/${TAG}_variants.txt
