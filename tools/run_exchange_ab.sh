set -x
mkdir -p gpurun_out
for ov in 0 1; do
RSC_OVERLAP_EXCHANGE=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 9 > gpurun_out/bench2_ov$ov.json 2> gpurun_out/bench2_ov$ov.err
tail -3 gpurun_out/bench2_ov$ov.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench2_ov$ov.json').read().strip().splitlines()[-1])
print('OV$ov', d['value'], d['ms_per_step'], d.get('ms_per_task'), d['config'].get('final_loss'), d.get('sustained',{}).get('value'))
PY
done
