#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/d_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/d_tests.log)"
grep -E "^FAILED" gpurun_out/d_tests.log | head
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 9 --no-cpu-baseline --sustained-s 0 > gpurun_out/d_bench$i.json 2> gpurun_out/d_bench$i.err
python - $i <<'PY'
import json, sys
d = json.loads(open('gpurun_out/d_bench%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
k = d['kernels']
print(round(d['value'], 2), {a: round(b, 2) for a, b in d['ms_per_task'].items()}, 'e2e', round(d['e2e']['value'], 2),
      'pm fwd/bwd ms', round(k['rsc_patch_merge_ln_fwd']['ms'], 2), k['rsc_patch_merge_ln_fwd'].get('big_frac'),
      round(k['rsc_patch_merge_ln_bwd']['ms'], 2), k['rsc_patch_merge_ln_bwd'].get('big_frac'))
PY
done
