python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for mode in payload peer; do echo "== FKGPU_MG=$mode"
FKGPU_MG=$mode FKGPU_MG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 3 --warmup 2 2>&1 | grep -E "metric|rror|\[mg\]" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['n_gpus'],'GPU', round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items()})
    except Exception as e: print(l[:700])
"
done
