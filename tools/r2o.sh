# round 2, call o (1 GPU): k_profile with the lookup out of line; L2 fetch granularity hint 32 / 64 (default) / 128 on the two random-access kernels
mkdir -p gpurun_out
( python -m pytest tests -q -m gpu -x -k "profiles_match_oracle or interleaved or relative_profiles_against or rem_" ) > gpurun_out/r2o_pytest_gpu.txt 2>&1
tail -2 gpurun_out/r2o_pytest_gpu.txt
run() { name=$1; shift; python bench.py --no-cpu --no-e2e --steps 3 "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2o_$name.err; }
run c4_def --config 4
FKGPU_L2_FETCH=32 run c4_f32 --config 4
FKGPU_L2_FETCH=128 run c4_f128 --config 4
FKGPU_L2_FETCH=32 run c2_f32
FKGPU_L2_FETCH=128 run c2_f128
python - <<'PY'
import json
for f in ["c4_def","c4_f32","c4_f128","c2_f32","c2_f128"]:
    try:
        d=json.loads([l for l in open("gpurun_out/r2o_%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), d.get("invariant_violations"))
    print("   ", d.get("all_stage_ms"))
PY
