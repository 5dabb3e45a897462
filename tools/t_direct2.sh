python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -5
for nd in 0 1; do echo "FKGPU_NODIRECT=$nd"; FKGPU_NODIRECT=$nd python bench.py --steps 4 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms'); print(d['e2e'])"; done
