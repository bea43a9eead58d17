#!/bin/bash
# round 2, session 18: one-barrier prologue, vote-sliced suffix in the warp dict-order replay
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_ab.py 0 2 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; echo "latency rc=$?"; cat gpurun_out/probe_latency.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; echo "phases rc=$?"; tail -17 gpurun_out/probe_phases.log
