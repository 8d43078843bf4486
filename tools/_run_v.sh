timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|Error" | tail -3
timeout 600 python bench.py --steps 30 --warmup 9 > gpurun_out/bench_r02_k.json 2> gpurun_out/bench_r02_k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_k.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['ms_per_task'].items()}, 'e2e', round(d['e2e']['value'],2), 'sus', round(d['sustained']['value'],2), d['config'].get('final_loss'), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline'].get('frac_per_launch_bound'), 'cpu', d['cpu_baseline']['value'])
for r in d['roofline_also']: print(' ', r['kernel'], r['bound'], round(r['frac'],3), r.get('frac_per_launch_bound'))
PY
