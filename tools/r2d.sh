# round 2, call d: oversize buckets -> record pipeline (spill), e2e block order, 20 Gbase multi-round batch
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/r2d_pytest_gpu.txt 2>&1
tail -25 gpurun_out/r2d_pytest_gpu.txt
python bench.py --steps 4 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2d_bench.err
python bench.py --steps 3 --no-cpu --e2e-order contig > gpurun_out/r2d_bench_contig.json 2> gpurun_out/r2d_bench_contig.err; echo "bench(contig) rc=$?"
FKGPU_VERBOSE=1 python bench.py --device-gen --genome-mbp 400 --steps 2 --warmup 1 > gpurun_out/r2d_bench_20g.json 2> gpurun_out/r2d_bench_20g.err; echo "bench(20G) rc=$?"; tail -4 gpurun_out/r2d_bench_20g.err
python - <<'PY'
import json
for f in ["r2d_bench","r2d_bench_contig","r2d_bench_20g"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), "| e2e", e.get("value"), e.get("ingest_ms_per_step"), e.get("finish_ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "rounds", d["config"].get("rounds"), d["config"].get("sorted_runs"), "kmers", d["config"]["kmers_per_gpu"])
    print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["roofline"]["frac"], d["gpu_launches"])
PY
