#!/bin/bash
# Round-2 hardware probes: TMA-staged 56-slot window tile + SWIZZLE_64B descriptors (tools/wmsa_probe.cu),
# each mode in its own process under a timeout.  Output: gpurun_out/probe2_*.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/wmsa_probe tools/wmsa_probe.cu || exit 1
for args in "plain 7 7" "plain 14 14" "plain -3 -3" "plain 0 17" "wrap"; do
  tag=$(echo "$args" | tr ' -' '_m')
  timeout 30 /tmp/wmsa_probe $args > gpurun_out/probe2_$tag.txt 2>&1
  echo "[$args] exit $?"; cat gpurun_out/probe2_$tag.txt
done
