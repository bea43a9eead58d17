#!/bin/bash
# round 2, session 14: early gather for up to 4 leading frees; source-level ncu capture of the step kernel, warm caches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_ab.py 0 4 2 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; echo "phases rc=$?"; tail -16 gpurun_out/probe_phases.log
NCU_REQ=4096 NCU_SPL=1 KVB200_TRACE_GRAPH=0 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:kvb_(step|decode) -c 3 --launch-skip 1500 \
  -o gpurun_out/r2_step_src -f python tools/ncu_target.py step > gpurun_out/ncu_step_src.log 2>&1; echo "ncu step src rc=$?"; tail -3 gpurun_out/ncu_step_src.log
ncu -i gpurun_out/r2_step_src.ncu-rep --page source --csv > gpurun_out/r2_step_src_source.csv 2> gpurun_out/r2_step_src_source.err; echo "source page rc=$?"
python tools/ncu_extract.py gpurun_out/r2_step_src.ncu-rep gpurun_out/r2_step_src.csv
gzip -f gpurun_out/r2_step_src_source.csv
ls -la gpurun_out/r2_step_src*; 
sz=$(stat -c %s gpurun_out/r2_step_src.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/r2_step_src.ncu-rep; fi
du -sh gpurun_out
