python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config1 or medium or repeats or forced" 2>&1 | tail -2
python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0})"
