# round 2, call n (1 GPU): profiles by hash lookup, written in output order and copied back in slices
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x -k "prof or relative or rem_ or direct_ingest or cli" ) > gpurun_out/r2n_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r2n_pytest_gpu.txt
( FKGPU_PROF=legacy python -m pytest tests -q -m gpu -x -k "profiles_match_oracle or interleaved or relative_profiles_against" ) > gpurun_out/r2n_pytest_legacy.txt 2>&1
tail -2 gpurun_out/r2n_pytest_legacy.txt
python bench.py --config 4 --steps 3 > gpurun_out/r2n_bench_config4.json 2> gpurun_out/r2n_bench_config4.err; echo "config 4 rc=$?"; tail -2 gpurun_out/r2n_bench_config4.err
ncu --set full --clock-control none --import-source on -k regex:'k_profile|k_hash_build' -c 3 -o gpurun_out/r2n_prof_profile -f \
    python bench.py --config 4 --no-cpu --no-e2e --steps 1 --warmup 0 > gpurun_out/r2n_prof_profile.log 2>&1
ls -la gpurun_out/r2n_prof_profile.ncu-rep
python - <<'PY'
import json
for f in ["r2n_bench_config4"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), d.get("step_wall_ms"), "| e2e", e.get("value"), e.get("ms_per_step"), "parity", d.get("parity_checked"), d.get("invariant_violations"))
    print("   ", d.get("all_stage_ms"), {k:e.get(k) for k in ("ingest_ms_per_step","finish_ms_per_step","profile_ms_per_step")})
PY
