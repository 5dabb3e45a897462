TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29611 tests/mgpu_worker.py 40 super-mer 2>&1 | grep -E "MGPU|rror" | head -3
FKGPU_MG=payload $TR --master-port 29612 tests/mgpu_worker.py 21 super-mer 2>&1 | grep -E "MGPU|rror" | head -3
$TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 1 2>&1 | grep -E "metric|rror" | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['n_gpus'],'GPU', round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', d['config']['parallelism'][:60])
    except Exception as e: print(l[:500])
"
