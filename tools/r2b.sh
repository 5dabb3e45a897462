# round 2, call b: the rewritten bucket kernel (k_bucket_count2) + wide entries (k 57..64 on the super-mer path)
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2b_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2b_pytest_gpu.txt
python bench.py --no-cpu --steps 4 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2b_bench.err
FKGPU_BC=old python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2b_bench_old.json 2> gpurun_out/r2b_bench_old.err; echo "bench(old) rc=$?"
for ts in 192 256; do FKGPU_TS=$ts python bench.py --no-cpu --no-e2e --steps 4 > gpurun_out/r2b_bench_ts$ts.json 2>/dev/null; done
for c in 3 5; do
  python bench.py --config $c --steps 3 > gpurun_out/r2b_bench_c$c.json 2> gpurun_out/r2b_bench_c$c.err; echo "config $c rc=$?"
  tail -2 gpurun_out/r2b_bench_c$c.err
done
python - <<'PY'
import json
for f in ["r2b_bench","r2b_bench_old","r2b_bench_ts192","r2b_bench_ts256","r2b_bench_c3","r2b_bench_c5"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), "| e2e", e.get("value"), "parity", d.get("parity_checked"), d.get("invariant_violations"))
    print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["roofline"]["frac"], d["gpu_launches"])
PY
# ncu: launch list + full capture of the bucket kernel on a 0.5 Gbase batch
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --genome-mbp 10 > gpurun_out/r2b_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bucket_count2 -s 1 -c 1 -o gpurun_out/r2b_prof -f \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --genome-mbp 10 > gpurun_out/r2b_prof.log 2>&1
ls -la gpurun_out/r2b_prof.ncu-rep
