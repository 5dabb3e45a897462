# round 2, call c: multi-round counts (mem_limit), chunked D2H behind the sort, warp-private bucket kernel A/B, 20 Gbase batch
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/r2c_pytest_gpu.txt 2>&1
tail -25 gpurun_out/r2c_pytest_gpu.txt
python bench.py --steps 4 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_bench.err
FKGPU_BC=warp python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2c_bench_warp.json 2> gpurun_out/r2c_bench_warp.err; echo "bench(warp) rc=$?"; tail -2 gpurun_out/r2c_bench_warp.err
for ts in 16 64; do FKGPU_BC=warp FKGPU_TS=$ts python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2c_bench_warp_ts$ts.json 2>/dev/null; done
for ts in 512 1024; do FKGPU_TS=$ts python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2c_bench_ts$ts.json 2>/dev/null; done
python bench.py --mem-limit-gb 24 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2c_bench_rounds.json 2> gpurun_out/r2c_bench_rounds.err; echo "bench(rounds) rc=$?"; tail -2 gpurun_out/r2c_bench_rounds.err
FKGPU_VERBOSE=1 python bench.py --device-gen --genome-mbp 400 --steps 2 --warmup 1 > gpurun_out/r2c_bench_20g.json 2> gpurun_out/r2c_bench_20g.err; echo "bench(20G) rc=$?"; tail -3 gpurun_out/r2c_bench_20g.err
python - <<'PY'
import json
for f in ["r2c_bench","r2c_bench_warp","r2c_bench_warp_ts16","r2c_bench_warp_ts64","r2c_bench_ts512","r2c_bench_ts1024","r2c_bench_rounds","r2c_bench_20g"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), "| e2e", e.get("value"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "rounds", d["config"].get("rounds"), d["config"].get("sorted_runs"))
    print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["roofline"]["frac"], d["gpu_launches"])
PY
