#!/bin/bash
# N independent single-GPU benchmarks at the same time on the N GPUs of one box (no process group): the spread of the
# per-GPU rates bounds what a synchronised data-parallel step can reach (it runs at the pace of the slowest GPU)
N=${1:-8}
mkdir -p gpurun_out
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i timeout 400 python bench.py --steps 60 --warmup 9 --no-cpu-baseline --sustained-s 3 > gpurun_out/replica_$i.json 2> gpurun_out/replica_$i.err &
done
wait
python - <<PY
import json
vals=[]
for i in range($N):
    try:
        d=json.loads(open('gpurun_out/replica_%d.json'%i).read().strip().splitlines()[-1])
        vals.append((i, round(d['value'],2), round(d['sustained']['value'],2), {k:round(v,2) for k,v in d['ms_per_task'].items()}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))
    except Exception as e:
        vals.append((i,'ERR',str(e)))
for v in vals: print(v)
PY
