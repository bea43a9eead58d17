#!/bin/bash
# round 2, session 16+: decode kernel, early-issued bitmap test, staged run masks, closing summary only
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_ab.py 0 8 2 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
timeout 300 python tools/probe_threads.py > gpurun_out/probe_t512.log 2>&1; echo "t512 rc=$?"; cat gpurun_out/probe_t512.log
timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; echo "latency rc=$?"; cat gpurun_out/probe_latency.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; echo "phases rc=$?"; tail -34 gpurun_out/probe_phases.log
