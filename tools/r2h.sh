# round 2, call h (1 GPU): group-level dedupe, range-normalised entry ordering, 8192-record partition tiles, rounds buffers kept
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/r2h_pytest_gpu.txt 2>&1
tail -6 gpurun_out/r2h_pytest_gpu.txt
python bench.py --steps 4 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2h_bench.err
python bench.py --config 3 --steps 3 --no-cpu --no-e2e > gpurun_out/r2h_bench_c3.json 2>/dev/null
python bench.py --config 5 --steps 3 --no-cpu --no-e2e > gpurun_out/r2h_bench_c5.json 2>/dev/null
python bench.py --coverage 5 --genome-mbp 400 --steps 3 --no-cpu --no-e2e > gpurun_out/r2h_bench_cov5.json 2>/dev/null
python bench.py --device-gen --genome-mbp 400 --steps 3 --warmup 1 > gpurun_out/r2h_bench_20g.json 2> gpurun_out/r2h_bench_20g.err; echo "20G rc=$?"
python - <<'PY'
import json
for f in ["r2h_bench","r2h_bench_c3","r2h_bench_c5","r2h_bench_cov5","r2h_bench_20g"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    c=d["config"]
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), "| e2e", e.get("value"), e.get("ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "sm", c.get("supermer_records"), "exp", c.get("supermers_expanded"), "split", c.get("split_classes"), "rounds", c.get("rounds"))
    print("   ", d.get("all_stage_ms"), d["roofline"]["frac"])
PY
