#!/bin/bash
# A/B of the round's scheduling changes on one B200 (one gpurun call, ~5 minutes): the new parity tests, then bench.py with
# each switch off in turn.  Outputs: gpurun_out/ab2_*.json + gpurun_out/ab2_tests.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest -x -q -m gpu tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_ops.py \
  -k "linear_pair or side_branch or graph_replay or patch_merge or msda_fused or step_engine" > gpurun_out/ab2_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/ab2_tests.log)"
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 9 --no-cpu-baseline --sustained-s 0 \
    > gpurun_out/ab2_$name.json 2> gpurun_out/ab2_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/ab2_%s.json' % name).read().strip().splitlines()[-1])
    print(name, round(d['value'], 2), round(d['ms_per_step'], 3), {k: round(v, 2) for k, v in d['ms_per_task'].items()},
          'e2e', round(d['e2e']['value'], 2), 'launches', d.get('gpu_launches'))
except Exception as e:
    print(name, 'FAILED', e, open('gpurun_out/ab2_%s.err' % name).read()[-1500:])
PY
}
run all_on RSC_X=0
run side_off RSC_SIDE_DW=0
run pair_off RSC_LINEAR_PAIR=0
run pm_v1 RSC_PATCH_MERGE_V1=1   # (at that commit v2 was the default; now: RSC_PATCH_MERGE_V2=1 opts in)
run all_on2 RSC_X=0
