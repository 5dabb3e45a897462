run() { python bench.py --steps 3 --warmup 2 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print(round(e['value'],2),'Gbases/s', round(e['ms_per_step'],1),'ms ingest',round(e['ingest_ms_per_step'],1),'finish',round(e['finish_ms_per_step'],1),'dev',round(e['finish_device_ms'],1), e['finish_stage_ms'])"; }
echo "direct 16thr"; run
echo "staged 16thr"; FKGPU_NODIRECT=1 run
echo "direct 8thr"; run --ingest-threads 8
echo "direct 32thr"; run --ingest-threads 32
echo "direct 16thr chunk 2MB"; FKGPU_CHUNK_BYTES=2097152 run
echo "direct 16thr chunk 32MB"; FKGPU_CHUNK_BYTES=33554432 run
