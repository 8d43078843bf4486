#!/bin/bash
# A/B of the gradient exchange at N GPUs (run under gpurun --gpus N): after-backward vs overlapped buckets
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 9 > gpurun_out/bench${N}_$tag.json 2> gpurun_out/bench${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(d['value'],2), round(d['ms_per_step'],3), d.get('ms_per_task'), 'sustained', round(d.get('sustained',{}).get('value',0),2), 'e2e', round(d['e2e']['value'],2))
except Exception as e:
    print('$tag ERR', e); print(open('gpurun_out/bench${N}_$tag.err').read()[-1500:])
PY
}
run ov0 RSC_OVERLAP_EXCHANGE=0
run ov1 RSC_OVERLAP_EXCHANGE=1
run ov1_cta8 RSC_OVERLAP_EXCHANGE=1 NCCL_MAX_CTAS=8
