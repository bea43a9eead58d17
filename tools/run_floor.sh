#!/bin/bash
# every variant in a fresh process; logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv > gpurun_out/floor_smi.txt
for v in zero nozero zero_kernel zero_ondemand zero_thr2 zero_bgkernel zero_sync; do
  timeout 300 ./tools/vmm_floor tools/events_decode4096.txt $v 3 128 > gpurun_out/floor_$v.log 2>&1
done
timeout 300 ./tools/vmm_floor tools/events_decode4096.txt zero 2 2 > gpurun_out/floor_zero_2MiB.log 2>&1
timeout 300 ./tools/vmm_floor tools/events_decode4096.txt nozero 2 2 > gpurun_out/floor_nozero_2MiB.log 2>&1
timeout 600 python tools/probe_e2e.py 100000 > gpurun_out/e2e_attr.log 2>&1
