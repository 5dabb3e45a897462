python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for v in 10 14; do echo "BC=$v"; FKGPU_BC=$v python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0})"; done
echo k21; python bench.py -k 21 --read-len 150 --steps 3 --warmup 2 --no-cpu --no-e2e --cutoff 4 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0})"
