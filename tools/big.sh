nvidia-smi --query-gpu=memory.used,memory.total --format=csv
FKGPU_VERBOSE=1 timeout 600 python bench.py --genome-mbp 88 --no-e2e --no-cpu --steps 1 --warmup 1 2>&1 | tail -4 | cut -c1-1500
