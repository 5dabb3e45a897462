# round 2, call k (8 GPUs): the in-library multi-GPU count at world 8: oracle parity tests, host program, bench line with parity
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x -k "in_library_equals_oracle or cli_multi_gpu" ) > gpurun_out/r2k_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2k_pytest_gpu.txt
FKGPU_MG_TIMING=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 \
      > gpurun_out/r2k_bench8.json 2> gpurun_out/r2k_bench8.err; echo "bench N=8 rc=$?"; tail -3 gpurun_out/r2k_bench8.err
FKGPU_MG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 1 --warmup 2 --no-cpu --no-e2e \
      > gpurun_out/r2k_bench8_tl.json 2> gpurun_out/r2k_bench8_tl.err; grep "fkgpu mg rank 0" gpurun_out/r2k_bench8_tl.err | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 4 --steps 3 --no-cpu --no-e2e \
      > gpurun_out/r2k_bench4.json 2> gpurun_out/r2k_bench4.err; echo "bench N=4 rc=$?"
python - <<'PY'
import json
for f in ["r2k_bench8","r2k_bench8_tl","r2k_bench4"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    c=d["config"]
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), "| e2e", e.get("value"), e.get("ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"), "sm", c.get("supermer_records"), "exp", c.get("supermers_expanded"), "split", c.get("split_classes"))
    print("   ", d.get("all_stage_ms"), d["roofline"]["frac"])
PY
