#!/bin/bash
# one GPU session: tests, bench (ours + reference arm), launch list of the bench, full captures of the kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
if [ "$1" = "ncu" ]; then
  NCU_SPL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_step.csv \
    python tools/ncu_target.py step > gpurun_out/ncu_list.log 2>&1
  NCU_REQ=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_step -c 1 -o gpurun_out/step_full_r1c -f \
    python tools/ncu_target.py step > gpurun_out/ncu_step.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_audit -c 1 --launch-skip 3 -o gpurun_out/audit_stress -f \
    python tools/probe_sweep.py > gpurun_out/ncu_audit.log 2>&1
fi
