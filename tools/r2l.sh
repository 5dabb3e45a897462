# round 2, call l (8 GPUs, short): host-link topology and what page placement does to concurrent table fetches
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; cat /proc/self/status | grep -i -E "allowed_list|Cap"; ls /sys/devices/system/node/ | head; } > gpurun_out/r2l_topo.txt 2>&1
timeout 200 gpurun_out/../tools/numa_probe > gpurun_out/r2l_probe.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/r2l_topo.txt | head -40
cat gpurun_out/r2l_probe.txt
