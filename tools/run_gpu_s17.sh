#!/bin/bash
# round 2, session 17: the driver's commands (bench N=1 K=20 W=5, reference arm), ncu launch list + full capture of the decode kernel
mkdir -p gpurun_out
ex() { python tools/ncu_extract.py gpurun_out/$1.ncu-rep gpurun_out/$1.csv; ncu -i gpurun_out/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
SECONDS=0
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$? wall=${SECONDS}s"; tail -c 600 gpurun_out/bench_ours.err
SECONDS=0
timeout 900 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference rc=$? wall=${SECONDS}s"; cat gpurun_out/bench_ref.json | cut -c1-400
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_ours.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms_per_step", d["ms_per_step"], "us/step", d["roofline"]["step_kernel"]["us_per_scheduler_step"])
print("e2e", d["e2e"]["value"], "retention", d["e2e"]["with_retention"]["value"], "no_zero", d["e2e"]["no_zero_page"].get("value"))
print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"].get("prealloc_off_value"), d["cpu_baseline"].get("bookkeeping_only_value"))
print("roofline", d["roofline"]["frac"], "micro", {k: round(v, 1) for k, v in d["micro_alloc_free"].items() if isinstance(v, float)})
print("per_call", d["e2e_per_call_api"]["value"], "mla", {k: v.get("ours_vs_reference") for k, v in d["mla_pin"].items() if isinstance(v, dict)})
print("two_model", json.dumps(d["two_model"])[:600])
PY
SECONDS=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench_v2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-two-model > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list rc=$? wall=${SECONDS}s"
python tools/summarize_launches.py gpurun_out/r2_launches_bench_v2.csv > gpurun_out/r2_launches_bench_v2.md 2>&1; cat gpurun_out/r2_launches_bench_v2.md
head -c 200000 gpurun_out/r2_launches_bench_v2.csv > gpurun_out/r2_launches_bench_v2_head.csv; rm -f gpurun_out/r2_launches_bench_v2.csv
NCU_REQ=4096 NCU_SPL=1 KVB200_TRACE_GRAPH=0 timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k 'regex:kvb_(step|decode)' -c 3 --launch-skip 1500 \
  -o gpurun_out/r2_decode_kernel -f python tools/ncu_target.py step > gpurun_out/ncu_decode.log 2>&1; echo "ncu decode rc=$?"; ex r2_decode_kernel
ls -la gpurun_out | tail -12
