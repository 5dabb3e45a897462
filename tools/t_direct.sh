python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for extra in "" "--cutoff 4"; do python bench.py --steps 3 --warmup 2 --no-cpu $extra 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0}, 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],1), 'table', d['config']['table_records'])"; done
