#!/bin/bash
# diagnostics of the multi-GPU overhead (run under gpurun --gpus N): what remains without the gradient exchange, NCCL stream priority
N=${1:-2}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 9 --sustained-s 0 > gpurun_out/diag${N}_$tag.json 2> gpurun_out/diag${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/diag${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d.get('ms_per_task').items()}, 'e2e', round(d['e2e']['value'],2))
except Exception as e:
    print('$tag ERR', e); print(open('gpurun_out/diag${N}_$tag.err').read()[-1500:])
PY
}
run ov1 RSC_OVERLAP_EXCHANGE=1
run skip RSC_SKIP_EXCHANGE=1
run ov1_hiprio RSC_OVERLAP_EXCHANGE=1 TORCH_NCCL_HIGH_PRIORITY=1
timeout 300 python bench.py --steps 30 --warmup 9 --sustained-s 0 --no-cpu-baseline > gpurun_out/diag1.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/diag1.json').read().strip().splitlines()[-1])
print('single', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d.get('ms_per_task').items()}, 'e2e', round(d['e2e']['value'],2))
"
