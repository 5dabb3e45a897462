# round-1 final evidence (one GPU): parity suite, full bench line, launch list + full ncu captures of the current pipeline
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r1c_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r1c_pytest_gpu.txt
FKGPU_VERBOSE=1 python bench.py > gpurun_out/r1c_bench_full.json 2> gpurun_out/r1c_bench_full.err
grep fkgpu gpurun_out/r1c_bench_full.err | tail -2
ARGS="--genome-mbp 10 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_" -c 60 --csv --log-file gpurun_out/r1c_launches.csv python bench.py $ARGS --steps 2 --warmup 1 > gpurun_out/r1c_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_(super|bucket_count|tilepart|refine|sortcount)" -c 8 -o gpurun_out/r1c_prof -f python bench.py $ARGS --steps 1 --warmup 0 > gpurun_out/r1c_prof.log 2>&1
python bench.py $ARGS --steps 3 --warmup 3 > gpurun_out/r1c_bench_small.json 2> gpurun_out/r1c_bench_small.err
python - <<'PY'
import json
for f in ("r1c_bench_full","r1c_bench_small"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),2), round(e.get("ms_per_step",0),1), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d.get("cpu_baseline",{}) and round(d["cpu_baseline"]["value"],3), d["roofline"]["frac"], d.get("clocks"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
