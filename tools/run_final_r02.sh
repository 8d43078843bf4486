#!/bin/bash
# end-of-round check on one B200: the whole -m gpu suite, smoke(), the default bench (with the CPU baseline and the sustained leg)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/final_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/final_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$? $(tail -2 gpurun_out/final_smoke.log | tr '\n' ' ')"
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(round(d['value'], 2), d['ms_per_step'], d['ms_per_task'], 'e2e', d['e2e']['value'], 'sustained', d.get('sustained', {}).get('value'),
      'cpu', d.get('cpu_baseline', {}).get('value'), 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
PY
