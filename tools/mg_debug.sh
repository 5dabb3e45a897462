export FKGPU_VERBOSE=1
for g in 2 20; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 1 --warmup 1 --genome-mbp $g 2>&1 | grep -E "fkgpu|metric" | cut -c1-700
done
