#!/bin/bash
# One-shot hardware probes (B200): builds tools/tc_probe_m64.cu / tc_probe_tma.cu and runs every mode in its own process
# (a faulting variant must not poison the others).  Output: gpurun_out/probe_m64_<mode>.txt
#   gpurun --timeout 300 -- 'bash tools/run_probes.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe_m64 tools/tc_probe_m64.cu || exit 1
for mode in m64 ld16x256 ld16x128 ld16x64 m64_lane16 m64_lane64 m128_lane64; do
  timeout 30 /tmp/tc_probe_m64 $mode > gpurun_out/probe_m64_$mode.txt 2>&1
  echo "$mode: exit $? ($(wc -l < gpurun_out/probe_m64_$mode.txt) lines)"
done
head -40 gpurun_out/probe_m64_m64.txt
# TMA box load of one window (4-D tensor map, SWIZZLE_64B) feeding the MMA directly
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe_tma tools/tc_probe_tma.cu || exit 1
for origin in "0 0" "14 14" "-3 -3"; do
  tag=$(echo "$origin" | tr ' -' '_m')
  timeout 30 /tmp/tc_probe_tma $origin > gpurun_out/probe_tma_$tag.txt 2>&1
  echo "tma origin ($origin): exit $? -- $(tail -1 gpurun_out/probe_tma_$tag.txt)"
done
