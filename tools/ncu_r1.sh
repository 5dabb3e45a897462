# ncu evidence for round 1 (one GPU, small batch so the 40x replays stay short)
KR="regex:k_(scan|scatter|colscan|refine|sortcount|compact|groups|lscan|fill|pack|tilepart|autorefine|suboff|mask)"
ARGS="--genome-mbp 10 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py $ARGS --steps 2 --warmup 1 > gpurun_out/launches_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_(scatter_tile|refine|sortcount)" -c 3 -o gpurun_out/prof_r1 -f python bench.py $ARGS --steps 1 --warmup 0 > gpurun_out/prof_r1.log 2>&1
python bench.py $ARGS --steps 3 --warmup 3 > gpurun_out/bench_small_r1.json 2> gpurun_out/bench_small_r1.err
tail -c 600 gpurun_out/bench_small_r1.json
