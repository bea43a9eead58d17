#!/bin/bash
# round 2, session 7: tests, bench N=1, then ncu: launch list of the bench, full captures of the copy kernel (2 KiB and
# 256 B cells) and of the step kernel (single request and decode step)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -6 gpurun_out/tests_gpu.log
timeout 1500 python bench.py --steps 3 --warmup 3 --ref-replays 1 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_ours.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-two-model > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
NCU_T=65536 NCU_CELL=2048 NCU_L=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_kv_copy -c 2 --launch-skip 4 \
  -o gpurun_out/r2_kv_copy_cell2048 -f python tools/ncu_target.py kv > gpurun_out/ncu_kv2048.log 2>&1; echo "ncu kv 2048 rc=$?"
NCU_T=65536 NCU_CELL=256 NCU_L=80 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_kv_copy -c 2 --launch-skip 4 \
  -o gpurun_out/r2_kv_copy_cell256 -f python tools/ncu_target.py kv > gpurun_out/ncu_kv256.log 2>&1; echo "ncu kv 256 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_step -c 2 --launch-skip 400 \
  -o gpurun_out/r2_step_single -f python tools/ncu_target.py single > gpurun_out/ncu_single.log 2>&1; echo "ncu single rc=$?"
NCU_REQ=4096 NCU_SPL=1 KVB200_TRACE_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kvb_step -c 3 --launch-skip 1500 \
  -o gpurun_out/r2_step_decode -f python tools/ncu_target.py step > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "with_retention", d["e2e"]["with_retention"]["value"], "no_zero", d["e2e"]["no_zero_page"].get("value"))
    print("two_model", json.dumps(d["two_model"], indent=1)[:3000])
    print("micro", d["micro_alloc_free"])
except Exception as e:
    print("no bench line:", e)
PY
