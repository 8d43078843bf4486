timeout 400 python bench.py --steps 30 --warmup 9 --no-cpu-baseline > gpurun_out/d2_single.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 9 > gpurun_out/d2_n2.json 2> gpurun_out/d2_n2.err
RSC_ASYNC_LOG_REDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 9 > gpurun_out/d2_n2_sync.json 2> gpurun_out/d2_n2_sync.err
python - <<'PY'
import json
for f in ('d2_single','d2_n2','d2_n2_sync'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['ms_per_task'].items()}, 'e2e', round(d['e2e']['value'],2), 'sus', round(d.get('sustained',{}).get('value',0),2), d['config'].get('final_loss'))
    except Exception as e:
        print(f,'ERR',e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
