#!/bin/bash
# round 2, session 15: one-barrier scans, CTA-wide execution of the warp alloc plan, 32-bit next_bkt; 256-thread build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_ab.py 0 2 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
KVB200_LIB=kvcached_b200/libkvb200_t256.so timeout 300 python tools/probe_threads.py > gpurun_out/probe_t256.log 2>&1; echo "t256 rc=$?"; cat gpurun_out/probe_t256.log
timeout 300 python tools/probe_threads.py > gpurun_out/probe_t512.log 2>&1; echo "t512 rc=$?"; cat gpurun_out/probe_t512.log
timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; echo "latency rc=$?"; cat gpurun_out/probe_latency.log
KVB200_LIB=kvcached_b200/libkvb200_t256.so timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency_t256.log 2>&1; echo "latency t256 rc=$?"; cat gpurun_out/probe_latency_t256.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; echo "phases rc=$?"; tail -16 gpurun_out/probe_phases.log
