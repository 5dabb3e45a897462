# round 2, call f: tuning sweep of the dedupe bucket kernel (bucket bits, group target, spill threshold) + new tests
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -k "relative or cli or library_world1 or stages_world1" ) > gpurun_out/r2f_pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2f_pytest_gpu.txt
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" python bench.py --no-cpu --no-e2e --steps 3 $EXTRA > gpurun_out/r2f_$name.json 2>/dev/null
}
EXTRA=""
for bb in 22 23 24; do for ts in 16 32; do run d_bb${bb}_ts${ts} FKGPU_BBITS=$bb FKGPU_TS=$ts; done; done
run d_bb23_ts24_big2048 FKGPU_BBITS=23 FKGPU_TS=24 FKGPU_BIG=2048
run d_bb22_ts32_big2048 FKGPU_BBITS=22 FKGPU_TS=32 FKGPU_BIG=2048
EXTRA="--config 3"
for bb in 22 24; do run c3_bb${bb}_big2048 FKGPU_BBITS=$bb FKGPU_BIG=2048; done
run c3_bb24_ts16_big1024 FKGPU_BBITS=24 FKGPU_TS=16 FKGPU_BIG=1024
EXTRA="--coverage 5 --genome-mbp 400"
for bb in 22 23 24; do run cov5_bb${bb} FKGPU_BBITS=$bb FKGPU_TS=16; done
EXTRA="--config 5"
run c5_bb22_big2048 FKGPU_BBITS=22 FKGPU_BIG=2048
run c5_bb24_ts16 FKGPU_BBITS=24 FKGPU_TS=16
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"no line",e); continue
    c=d["config"]
    print(f.split("r2f_")[1][:-5].ljust(24), round(d["value"],2), "Gb/s dev", round(d["device_ms_per_step"],1), "ms", {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, "sm", c.get("supermer_records"), "exp", c.get("supermers_expanded"), "split", c.get("split_classes"), "spill", c.get("spilled_kmers"), d.get("invariant_violations"))
PY
