KR="regex:k_"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 40 --csv --log-file gpurun_out/launches_super.csv python bench.py --genome-mbp 10 --no-cpu --no-e2e --steps 1 --warmup 0 > gpurun_out/launches_super.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_super.csv') if l.startswith('"')))
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:70].ljust(70), r[iv])
PY
