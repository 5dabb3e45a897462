# round-1 (second session) GPU evidence: parity suite, full bench line, launch list + full ncu captures of the super-mer path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r1b_smi.txt
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r1b_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r1b_pytest_gpu.txt
python bench.py > gpurun_out/r1b_bench_full.json 2> gpurun_out/r1b_bench_full.err
python bench.py --cutoff 0 --no-cpu --no-e2e > gpurun_out/r1b_bench_hist.json 2> gpurun_out/r1b_bench_hist.err
ARGS="--genome-mbp 10 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_" -c 60 --csv --log-file gpurun_out/r1b_launches.csv python bench.py $ARGS --steps 2 --warmup 1 > gpurun_out/r1b_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_(super|bucket_count|tilepart|refine|sortcount|compact)" -c 9 -o gpurun_out/r1b_prof -f python bench.py $ARGS --steps 1 --warmup 0 > gpurun_out/r1b_prof.log 2>&1
python bench.py $ARGS --steps 3 --warmup 3 > gpurun_out/r1b_bench_small.json 2> gpurun_out/r1b_bench_small.err
python - <<'PY'
import json
for f in ("r1b_bench_full","r1b_bench_hist","r1b_bench_small"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms", "e2e", d.get("e2e"), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d.get("cpu_baseline"), d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
