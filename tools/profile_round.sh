#!/bin/bash
# Refresh the profiling evidence of a round on one B200 (run under gpurun, ~8 minutes):
#   1. ncu launch list restricted to bench.py's timed region (NVTX range "timed_region"), 3 steps = one cls/det/seg cycle
#   2. ncu --set full of the roofline kernels at their stage-0 size (tools/run_wmsa_once.py), of the fused element-wise /
#      ms_deform_attn kernels and of the decoder attention kernels
# Outputs land in gpurun_out/; summarise into profiles/rNN_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --nvtx --nvtx-include "timed_region/" --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_timed_region.csv python bench.py --steps 3 --warmup 9 --no-cpu-baseline --sustained-s 0 > gpurun_out/launches_bench.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches_timed_region.csv) lines"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wmsa -c 4 -o gpurun_out/wmsa_stage0 -f \
  python tools/run_wmsa_once.py > gpurun_out/ncu_wmsa.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd|attn_bwd|m2f_mask" -c 8 -o gpurun_out/attn -f \
  python tools/run_attn_once.py > gpurun_out/ncu_attn.log 2>&1
if [ "${PROFILE_FUSED:-0}" = 1 ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bias_act|add_ln|msda_fused" -c 8 -o gpurun_out/fused -f \
  python tools/run_fused_once.py > gpurun_out/ncu_fused.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep 2>/dev/null
