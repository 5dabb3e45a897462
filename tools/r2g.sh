# round 2, call g (2 GPUs): multi-GPU paths after the record-format / dedupe changes: Python-driven and inside the C library
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -k "multi or fastmerge" ) > gpurun_out/r2g_pytest_gpu.txt 2>&1
tail -12 gpurun_out/r2g_pytest_gpu.txt
for impl in c py; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --mg-impl $impl \
      > gpurun_out/r2g_bench2_$impl.json 2> gpurun_out/r2g_bench2_$impl.err; echo "bench N=2 ($impl) rc=$?"; tail -3 gpurun_out/r2g_bench2_$impl.err
done
python bench.py --steps 4 --no-cpu --no-e2e > gpurun_out/r2g_bench1.json 2>/dev/null
python - <<'PY'
import json
for f in ["r2g_bench1","r2g_bench2_c","r2g_bench2_py"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), "| e2e", e.get("value"), e.get("ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"), (d.get("parity") or {}).get("via"))
    print("   ", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["roofline"]["frac"], d["gpu_launches"])
PY
