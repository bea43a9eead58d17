#!/bin/bash
# round 2, session 9: ncu evidence.  Reports are extracted to CSV on the box and deleted (gpurun_out must stay small).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_manager_parity.py -m gpu -x -q > gpurun_out/tests_parity.log 2>&1; echo "parity pytest rc=$?"; tail -3 gpurun_out/tests_parity.log
ex() { python tools/ncu_extract.py gpurun_out/$1.ncu-rep gpurun_out/$1.csv; ncu -i gpurun_out/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-two-model > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.md 2>&1; cat gpurun_out/r2_launches_bench.md
head -c 300000 gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench_head.csv; rm -f gpurun_out/r2_launches_bench.csv
NCU_T=65536 NCU_CELL=2048 NCU_L=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_kv_copy -c 2 --launch-skip 4 \
  -o gpurun_out/r2_kv_copy_cell2048 -f python tools/ncu_target.py kv > gpurun_out/ncu_kv2048.log 2>&1; echo "ncu kv 2048 rc=$?"; ex r2_kv_copy_cell2048
NCU_T=65536 NCU_CELL=256 NCU_L=80 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_kv_copy -c 2 --launch-skip 4 \
  -o gpurun_out/r2_kv_copy_cell256 -f python tools/ncu_target.py kv > gpurun_out/ncu_kv256.log 2>&1; echo "ncu kv 256 rc=$?"; ex r2_kv_copy_cell256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_step -c 2 --launch-skip 400 \
  -o gpurun_out/r2_step_single -f python tools/ncu_target.py single > gpurun_out/ncu_single.log 2>&1; echo "ncu single rc=$?"; ex r2_step_single
NCU_REQ=4096 NCU_SPL=1 KVB200_TRACE_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_step -c 3 --launch-skip 1500 \
  -o gpurun_out/r2_step_decode -f python tools/ncu_target.py step > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"; ex r2_step_decode
ls -la gpurun_out | tail -20; du -sh gpurun_out
