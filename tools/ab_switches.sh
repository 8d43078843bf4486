#!/bin/bash
# A/B of the opt-in kernel variants on one B200 (same box, back to back):
#   gpurun --timeout 900 -- 'bash tools/ab_switches.sh'
# 1. parity of the variants (the experimental tests + the regular add_ln / bias tests with the switches on)
# 2. kernel micro-benchmarks (tools/kbench.py covers both GELU variants)
# 3. the co-training bench with each switch
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity with the switches on"
RSC_TEST_EXPERIMENTAL=1 RSC_GELU_SIG=1 RSC_ADD_LN_LEAN=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x \
  -k "bias_gelu or add_ln or logistic" 2>&1 | tail -3
RSC_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_golden_reference.py -m gpu -q -k inference 2>&1 | tail -2
echo "== kbench (default switches)"
KB_B=4 timeout 300 python tools/kbench.py 2>/dev/null | grep -E "bias_act|add_ln" | tee gpurun_out/ab_kbench_default.jsonl
echo "== kbench (RSC_ADD_LN_LEAN=1)"
RSC_ADD_LN_LEAN=1 KB_B=4 timeout 300 python tools/kbench.py 2>/dev/null | grep -E "add_ln" | tee gpurun_out/ab_kbench_lean.jsonl
for sw in "" "RSC_GELU_SIG=1" "RSC_ADD_LN_LEAN=1" "RSC_COMPACT_ATTN_MASK=1" "RSC_CONST_ATTN_BIAS=1" "RSC_GELU_SIG=1 RSC_ADD_LN_LEAN=1 RSC_COMPACT_ATTN_MASK=1 RSC_CONST_ATTN_BIAS=1"; do
  tag=$(echo "${sw:-default}" | tr ' =' '__')
  echo "== bench [$sw]"
  env $sw timeout 240 python bench.py --steps 15 --warmup 6 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab_bench_$tag.json
  python - <<PY
import json
d = json.load(open('gpurun_out/ab_bench_$tag.json'))
print('$tag', 'value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'ms/task', {k: round(v, 2) for k, v in d.get('ms_per_task', {}).items()})
PY
done
