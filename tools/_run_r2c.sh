set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 30 --warmup 9 > gpurun_out/bench_r02_i.json 2> gpurun_out/bench_r02_i.err
for c in configs/multi/det_only_swin-s_800.py configs/seg/upernet_swin-b_512_potsdam.py; do
n=$(basename $c .py)
timeout 600 python bench.py --config $c --steps 20 --warmup 6 --no-cpu-baseline > gpurun_out/bench_i_$n.json 2> gpurun_out/bench_i_$n.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02_i.json')+glob.glob('gpurun_out/bench_i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e']['value'], d.get('ms_per_task'), d.get('sustained',{}).get('value'), d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
