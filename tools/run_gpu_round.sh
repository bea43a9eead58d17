#!/bin/bash
# one GPU session: tests, bench (ours + reference arm), launch list of the bench, full captures of the two kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  NCU_T=65536 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_kv_copy -c 2 -o gpurun_out/kv_T65536 -f \
    python tools/ncu_target.py kv > gpurun_out/ncu_kv65536.log 2>&1
fi
