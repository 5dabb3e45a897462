# round 2, call i (2 GPUs): multi-GPU after compact payload / local positions / group-level dedupe
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x -k "multi or cli_multi" ) > gpurun_out/r2i_pytest_gpu.txt 2>&1
tail -6 gpurun_out/r2i_pytest_gpu.txt
python bench.py --steps 5 --no-cpu --no-e2e > gpurun_out/r2i_bench1.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 \
      > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err; echo "bench N=2 rc=$?"; tail -3 gpurun_out/r2i_bench2.err
FKGPU_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 3 --config 5 --no-e2e \
      > gpurun_out/r2i_bench2_c5.json 2> gpurun_out/r2i_bench2_c5.err; echo "bench N=2 c5 rc=$?"; tail -3 gpurun_out/r2i_bench2_c5.err
python - <<'PY'
import json
for f in ["r2i_bench1","r2i_bench2","r2i_bench2_c5"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    c=d["config"]
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), "| e2e", e.get("value"), e.get("ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "sm", c.get("supermer_records"), "exp", c.get("supermers_expanded"), "split", c.get("split_classes"), d.get("clocks"))
    print("   ", d.get("all_stage_ms"), d["roofline"]["frac"])
PY
