# round 2, call q (1 GPU, last): profile lookups with / without the L2::64B hint on the 256-bit load
mkdir -p gpurun_out
( FKGPU_PROF_LTC=64 python -m pytest tests -q -m gpu -x -k "profiles_match_oracle or interleaved or relative_profiles_against" ) > gpurun_out/r2q_pytest_gpu.txt 2>&1
tail -2 gpurun_out/r2q_pytest_gpu.txt
FKGPU_PROF_LTC=64 python bench.py --config 4 --no-cpu --no-e2e --steps 3 > gpurun_out/r2q_c4_ltc64.json 2> gpurun_out/r2q_c4_ltc64.err; echo "rc=$?"
FKGPU_PROF_LTC=64 ncu --set full --clock-control none --import-source on -k regex:'k_profile' -c 1 -o gpurun_out/r2q_prof_ltc64 -f \
    python bench.py --config 4 --no-cpu --no-e2e --steps 1 --warmup 0 > gpurun_out/r2q_prof.log 2>&1
python - <<'PY'
import json
for f in ["r2q_c4_ltc64"]:
    d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), d.get("invariant_violations"))
    print("   ", d.get("all_stage_ms"))
PY
