python tools/e2e_probe.py 2>&1 | grep -v Warn | tail -8
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "bilinear" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_uper_head.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
for c in configs/seg/upernet_swin-b_512_potsdam.py; do
n=$(basename $c .py)
timeout 600 python bench.py --config $c --steps 20 --warmup 6 --no-cpu-baseline > gpurun_out/bench_l_$n.json 2> gpurun_out/bench_l_$n.err
done
timeout 600 python bench.py --steps 30 --warmup 9 --no-cpu-baseline --sustained-s 0 > gpurun_out/bench_r02_l.json 2> gpurun_out/bench_r02_l.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_l_*.json'))+['gpurun_out/bench_r02_l.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), 'e2e', round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['ms_per_task'].items()}, d['gpu_launches'])
        ks=d['kernels']
        print('  ', [(k, round(v['ms'],1)) for k,v in sorted(ks.items(), key=lambda kv:-kv[1]['ms'])[:8]])
    except Exception as e: print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
