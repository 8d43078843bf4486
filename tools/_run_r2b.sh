set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/abench.py r02_abench 2>&1 | tail -8
for ab in "1 1" "0 0"; do set -- $ab
RSC_OWN_ATTN=$1 RSC_MASK_BITS=$2 timeout 600 python bench.py --steps 30 --warmup 9 > gpurun_out/bench_attn_$1.json 2> gpurun_out/bench_attn_$1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_attn_$1.json').read().strip().splitlines()[-1])
print('ATTN$1', d['value'], d['e2e']['value'], d.get('ms_per_task'), d['config'].get('final_loss'), d.get('sustained',{}).get('value'), d['gpu_launches'])
PY
done
