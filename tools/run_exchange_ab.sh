#!/bin/bash
# A/B of the gradient exchange at N GPUs (run under gpurun --gpus N): NCCL after backward / NCCL overlapped buckets /
# in-switch (NVLS multimem) overlapped buckets
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
    print('$tag', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d.get('ms_per_task').items()}, 'sustained', round(d.get('sustained',{}).get('value',0),2), 'e2e', round(d['e2e']['value'],2), 'loss', d['config'].get('final_loss'))
except Exception as e:
    print('$tag ERR', e); print(open('gpurun_out/bench${N}_$tag.err').read()[-2500:])
PY
}
for t in ${2:-ov0 ov1 nvls}; do
  case $t in
    ov0) run ov0 RSC_OVERLAP_EXCHANGE=0 ;;
    ov1) run ov1 RSC_OVERLAP_EXCHANGE=1 ;;
    nvls) run nvls RSC_OVERLAP_EXCHANGE=1 RSC_NVLS=1 ;;
    nvls0) run nvls0 RSC_OVERLAP_EXCHANGE=0 RSC_NVLS=1 ;;
  esac
done
