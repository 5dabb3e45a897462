python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu -k "forced or world1 or medium or repeats" 2>&1 | tail -2
for v in "" _v128 ""; do echo "lib$v"; FKGPU_VERBOSE=1 FKGPU_LIB=$PWD/fastk_b200/lib/libfastk_gpu$v.so python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e 2>&1 | grep -E "overflow|metric" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('[fkgpu]'): print(l.strip()[-60:]); continue
    d=json.loads(l); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0})"; done
