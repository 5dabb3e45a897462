python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for n in 4; do
FKGPU_MG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $n --steps 3 --warmup 2 2>&1 | grep -E "metric|rror|\[mg\]" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['n_gpus'],'GPU', round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms dev', round(d['device_ms_per_step'],1), d['config']['pipeline'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})
    except Exception as e: print(l[:700])
"
done
