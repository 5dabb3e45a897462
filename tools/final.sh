mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r1d_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r1d_pytest_gpu.txt
python bench.py > gpurun_out/r1d_bench_full.json 2> gpurun_out/r1d_bench_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r1d_bench_full.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms | e2e", round(e["value"],2), round(e["ms_per_step"],1), "ingest", round(e["ingest_ms_per_step"],1), "finish", round(e["finish_ms_per_step"],1), e["finish_stage_ms"])
print({k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["cpu_baseline"]["value"], d["roofline"]["frac"], d["clocks"], d["gpu_launches"])
PY
