for v in "11 22" "10 22" "9 22" "11 20" "10 20" "9 19" "8 18"; do set -- $v; echo "SP1=$1 SBB=$2"; FKGPU_SP1=$1 FKGPU_SBB=$2 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --cutoff 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0})"; done
