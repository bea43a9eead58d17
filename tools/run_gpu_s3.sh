#!/bin/bash
# round 2, session 3: phase profile, fast path latency, copy kernel variants, tests
mkdir -p gpurun_out
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; cat gpurun_out/probe_phases.log
timeout 300 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; tail -16 gpurun_out/probe_latency.log
for mlp in 4 8; do for ord in s c; do
  echo "== copy kernel: KVB200_COPY_MLP=$mlp KVB200_COPY_ORDER=$ord"
  KVB200_COPY_MLP=$mlp KVB200_COPY_ORDER=$ord timeout 300 python tools/probe_scatter.py 2>&1 | tail -6
done; done > gpurun_out/probe_scatter.log 2>&1
cat gpurun_out/probe_scatter.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
