#!/bin/bash
# (record of a measurement: RSC_ADDLN_UNROLL existed only at the commit this was run on -- the two-row-group variant lost
# 1.3 % and was reverted, profiles/r02_ab_round2b.log)
# second A/B of the round on one B200: new parity tests, add+LayerNorm micro-benchmark with one / two row groups per warp step,
# bench.py with the switch both ways.  Outputs: gpurun_out/ab3_*
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest -q -m gpu tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_ops.py \
  -k "linear_pair or graph_replay or patch_merge or msda_fused or step_engine or add_ln" > gpurun_out/ab3_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/ab3_tests.log)"
RSC_ADDLN_UNROLL=1 timeout 200 python tools/lnbench.py ab3_lnbench_u1 > /dev/null 2>&1
RSC_ADDLN_UNROLL=2 timeout 200 python tools/lnbench.py ab3_lnbench_u2 > /dev/null 2>&1
python - <<'PY'
import json
for u in (1, 2):
    try:
        for l in open('gpurun_out/ab3_lnbench_u%d.jsonl' % u):
            d = json.loads(l)
            print('U%d' % u, d['name'], 'add_ln fwd', d['add_ln_fwd_us'], 'bwd', d['add_ln_bwd_us'])
    except Exception as e:
        print('lnbench', u, 'FAILED', e)
PY
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 9 --no-cpu-baseline --sustained-s 0 \
    > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/ab3_%s.json' % name).read().strip().splitlines()[-1])
    print(name, round(d['value'], 2), round(d['ms_per_step'], 3), {k: round(v, 2) for k, v in d['ms_per_task'].items()},
          'e2e', round(d['e2e']['value'], 2), 'launches', d.get('gpu_launches'))
except Exception as e:
    print(name, 'FAILED', e, open('gpurun_out/ab3_%s.err' % name).read()[-1500:])
PY
}
run u2 RSC_ADDLN_UNROLL=2
run u1 RSC_ADDLN_UNROLL=1
run u2b RSC_ADDLN_UNROLL=2
run u1b RSC_ADDLN_UNROLL=1
