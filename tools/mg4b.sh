nvidia-smi topo -m 2>&1 | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29601 tools/p2p_probe.py 2>&1 | grep -E "peer|ranks|rror" | head
FKGPU_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 4 --steps 2 --warmup 1 --genome-mbp 10 2>&1 | grep -E "metric|rror|fkgpu" | tail -3 | cut -c1-400
