# round 2, call j (2 GPUs): phase timeline of the in-library multi-GPU count
mkdir -p gpurun_out
FKGPU_MG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu --no-e2e \
      > gpurun_out/r2j_bench2.json 2> gpurun_out/r2j_bench2.err; echo "rc=$?"
grep "fkgpu mg rank 0" gpurun_out/r2j_bench2.err | tail -3
grep "fkgpu mg rank 1" gpurun_out/r2j_bench2.err | tail -1
