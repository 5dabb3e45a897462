python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for diag in "" 1; do
echo "== FKGPU_DIAG_LOCALSEQ=$diag"
FKGPU_DIAG_LOCALSEQ=$diag FKGPU_VERBOSE=1 FKGPU_MG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 4 --steps 2 --warmup 1 2>&1 | grep -E "rror|super_count|ms_per_step|\[mg\]" | tail -4 | cut -c1-330
done
