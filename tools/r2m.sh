# round 2, call m (1 GPU): e2e arm with C producer threads vs Python threads, thread counts
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py --no-cpu --steps 5 "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/r2m_$name.err; }
run py16 --e2e-feeder py
run c16 --e2e-feeder c
run c8 --e2e-feeder c --ingest-threads 8
run c32 --e2e-feeder c --ingest-threads 32
run c4 --e2e-feeder c --ingest-threads 4
python - <<'PY'
import json
for f in ["py16","c16","c8","c32","c4"]:
    try:
        d=json.loads([l for l in open("gpurun_out/r2m_%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"no line",e); continue
    e=d.get("e2e") or {}
    print(f, round(d["value"],2), "Gbases/s", round(d["ms_per_step"],1), "ms dev", round(d["device_ms_per_step"],1), "| e2e", round(e.get("value"),2), round(e.get("ms_per_step"),1), "ingest", round(e.get("ingest_ms_per_step"),1), "finish", round(e.get("finish_ms_per_step"),1), d.get("invariant_violations"))
PY
