python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k world1 2>&1 | tail -2
for v in "" _v1 _v2 _v4 _v7 ""; do echo "lib$v"; FKGPU_LIB=$PWD/fastk_b200/lib/libfastk_gpu$v.so python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'Gbases/s', round(d['ms_per_step'],1),'ms', {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0})"; done
