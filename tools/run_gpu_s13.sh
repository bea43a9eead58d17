#!/bin/bash
# round 2, session 13: programmatic dependent launch + early ref gather: full GPU suite, same-box A/B, latency and phase probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_ab.py 0 2 4 6 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
KVB200_TUNE=0 timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency_pdl.log 2>&1; echo "latency (PDL) rc=$?"; cat gpurun_out/probe_latency_pdl.log
KVB200_TUNE=2 timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency_nopdl.log 2>&1; echo "latency (no PDL) rc=$?"; cat gpurun_out/probe_latency_nopdl.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; echo "phases rc=$?"; tail -34 gpurun_out/probe_phases.log
