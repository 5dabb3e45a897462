# round 2, call a: parity at scale + new bench contract + the other configs (1 GPU)
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2a_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2a_pytest_gpu.txt
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
for c in 3 5; do
  python bench.py --config $c --steps 3 > gpurun_out/r2a_bench_c$c.json 2> gpurun_out/r2a_bench_c$c.err; echo "config $c rc=$?"
  tail -2 gpurun_out/r2a_bench_c$c.err
done
python - <<'PY'
import json
for f in ["r2a_bench","r2a_bench_c3","r2a_bench_c5","r2a_bench_ref"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms | e2e", e.get("value"), "parity", d.get("parity_checked"), d.get("invariant_violations"))
    if "roofline" in d:
        print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, (d.get("cpu_baseline") or {}).get("value"), d["roofline"]["frac"], d["gpu_launches"])
PY
nproc; free -g | head -2
