#!/bin/bash
# A/B of the branch-free two-point ms_deform_attn backward on one B200
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest -q -m gpu tests/test_gpu_ops.py tests/test_golden_reference.py -k "msda or deform or dino or encoder" > gpurun_out/e_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/e_tests.log)"
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 9 --no-cpu-baseline --sustained-s 0 > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/e_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
k = d['kernels']['rsc_msda_fused_bwd']
print(sys.argv[1], round(d['value'], 2), {a: round(b, 2) for a, b in d['ms_per_task'].items()}, 'e2e', round(d['e2e']['value'], 2),
      'msda bwd ms', round(k['ms'], 2), 'launches', k['launches'])
PY
}
run bf1 RSC_MSDA_BWD_BF=1
run bf0 RSC_MSDA_BWD_BF=0
run bf1b RSC_MSDA_BWD_BF=1
run bf0b RSC_MSDA_BWD_BF=0
