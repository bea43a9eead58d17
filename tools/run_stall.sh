#!/bin/bash
mkdir -p gpurun_out
{ nproc; cat /proc/cpuinfo | grep "model name" | sort | uniq -c; nvidia-smi; nvidia-smi --query-compute-apps=pid,name,used_memory --format=csv; ps aux --sort=-%cpu | head -15; } > gpurun_out/stall_env.txt 2>&1
./tools/vmm_stall 12 64 128 > gpurun_out/stall_h64.log 2>&1
./tools/vmm_stall 12 880 128 > gpurun_out/stall_h880.log 2>&1
./tools/vmm_stall 12 64 2 > gpurun_out/stall_h64_2m.log 2>&1
taskset -c 2 ./tools/vmm_stall 12 64 128 > gpurun_out/stall_h64_pinned.log 2>&1
ps aux --sort=-%cpu | head -8 >> gpurun_out/stall_env.txt
