ARGS="--genome-mbp 10 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:k_(super|bucket_count)" -c 2 -o gpurun_out/r1c_prof -f python bench.py $ARGS --steps 1 --warmup 0 > gpurun_out/r1c_prof.log 2>&1
tail -3 gpurun_out/r1c_prof.log
