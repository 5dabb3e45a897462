# round 2, call e: sequence-defined super-mers + duplicate super-mers counted once (warp-private bucket kernel), spill fix
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/r2e_pytest_gpu.txt 2>&1
tail -25 gpurun_out/r2e_pytest_gpu.txt
python bench.py --steps 4 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2e_bench.err
FKGPU_BC=cta python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2e_bench_cta.json 2> gpurun_out/r2e_bench_cta.err; echo "bench(cta) rc=$?"
for ts in 24 64 128; do FKGPU_TS=$ts python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2e_bench_ts$ts.json 2>/dev/null; done
for c in 3 5; do
  python bench.py --config $c --steps 3 > gpurun_out/r2e_bench_c$c.json 2> gpurun_out/r2e_bench_c$c.err; echo "config $c rc=$?"; tail -2 gpurun_out/r2e_bench_c$c.err
done
python bench.py --coverage 5 --genome-mbp 400 --steps 3 --no-cpu --no-e2e > gpurun_out/r2e_bench_cov5.json 2>/dev/null; echo "cov5 rc=$?"
python - <<'PY'
import json
for f in ["r2e_bench","r2e_bench_cta","r2e_bench_ts24","r2e_bench_ts64","r2e_bench_ts128","r2e_bench_c3","r2e_bench_c5","r2e_bench_cov5"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    c=d["config"]
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), "| e2e", e.get("value"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "supermers", c.get("supermer_records"), "expanded", c.get("supermers_expanded"), "split", c.get("split_classes"), "spill", c.get("spilled_kmers"))
    print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["roofline"]["frac"], d["gpu_launches"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --genome-mbp 10 > gpurun_out/r2e_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bucket_count3 -s 1 -c 1 -o gpurun_out/r2e_prof -f \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --genome-mbp 10 > gpurun_out/r2e_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_super -s 1 -c 1 -o gpurun_out/r2e_prof_super -f \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --genome-mbp 10 > gpurun_out/r2e_prof_super.log 2>&1
ls -la gpurun_out/r2e_prof*.ncu-rep
