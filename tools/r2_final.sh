# round 2, evidence run (1 GPU): full GPU test log, bench lines of every config, ncu launch list + full captures, sanitizer
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/r2_pytest_gpu.txt 2>&1
tail -6 gpurun_out/r2_pytest_gpu.txt
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err
for c in 3 4 5; do
  python bench.py --config $c --steps 3 > gpurun_out/r2_bench_config$c.json 2> gpurun_out/r2_bench_config$c.err; echo "config $c rc=$?"; tail -2 gpurun_out/r2_bench_config$c.err
done
python bench.py --device-gen --genome-mbp 400 --steps 2 --warmup 1 > gpurun_out/r2_bench_20gbases.json 2> gpurun_out/r2_bench_20gbases.err; echo "20G rc=$?"
python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --genome-mbp 10 > gpurun_out/r2_bench_small.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --genome-mbp 10 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_super|k_tilepart|k_refine|k_bucket_count3|k_sortcount' -s 12 -c 14 -o gpurun_out/r2_prof -f \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --genome-mbp 10 > gpurun_out/r2_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_profile|k_gather_profile' -c 2 -o gpurun_out/r2_prof_profile -f \
    python bench.py --config 4 --no-cpu --no-e2e --steps 1 --warmup 0 --genome-mbp 10 > gpurun_out/r2_prof_profile.log 2>&1
ls -la gpurun_out/r2_prof*.ncu-rep
( compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_config1_1k_reads or test_edge_cases or multi_round" ) > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
( compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_config1_1k_reads and 40" ) > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2_sanitizer_memcheck.txt; tail -4 gpurun_out/r2_sanitizer_racecheck.txt
python - <<'PY'
import json
for f in ["r2_bench","r2_bench_config3","r2_bench_config4","r2_bench_config5","r2_bench_20gbases","r2_bench_small"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), "| e2e", e.get("value"), e.get("ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["roofline"]["frac"], d["gpu_launches"], d["config"].get("rounds"))
PY
