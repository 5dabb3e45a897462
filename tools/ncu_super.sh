ARGS="--genome-mbp 10 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:k_(super|bucket_count)" -c 2 -o gpurun_out/prof_super -f python bench.py $ARGS --steps 1 --warmup 0 > gpurun_out/prof_super.log 2>&1
tail -c 300 gpurun_out/prof_super.log
