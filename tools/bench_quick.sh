python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for extra in "" "--cutoff 0"; do
FKGPU_VERBOSE=1 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $extra 2>&1 | grep -E "fkgpu|metric" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('[fkgpu]'): print(l.strip()); continue
    d=json.loads(l); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items()})
"
done
