timeout 600 python bench.py --steps 30 --warmup 9 > gpurun_out/bench_r02_m.json 2> gpurun_out/bench_r02_m.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_m.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['ms_per_task'].items()}, 'e2e', round(d['e2e']['value'],2), 'sus', round(d['sustained']['value'],2), d['config'].get('final_loss'), 'cpu', d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
