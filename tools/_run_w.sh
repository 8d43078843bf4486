python tools/e2e_probe.py 2>&1 | grep -v Warn | tail -20
for sd in 1 0; do
RSC_SIDE_DW=$sd timeout 400 python bench.py --steps 30 --warmup 9 --no-cpu-baseline --sustained-s 0 > gpurun_out/bench_side$sd.json 2>gpurun_out/bench_side$sd.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_side$sd.json').read().strip().splitlines()[-1])
    print('SIDE$sd', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['ms_per_task'].items()}, 'e2e', round(d['e2e']['value'],2), d['config'].get('final_loss'), d['cuda_graph_capture_failures'])
except Exception as e:
    print('SIDE$sd ERR', e); print(open('gpurun_out/bench_side$sd.err').read()[-3000:])
PY
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_golden_reference.py -m gpu -x -q 2>&1 | tail -3
