#!/bin/bash
# round 2, session 20: constant first rehash phase skipped; finer phase marks; smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python tools/probe_ab.py 0 2 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; echo "phases rc=$?"; tail -20 gpurun_out/probe_phases.log
