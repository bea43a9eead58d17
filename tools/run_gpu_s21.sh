#!/bin/bash
# round 2, session 21: final kernel (redundant barriers around free() removed, alloc counters on the last warp): full suite, smoke,
# A/B, then the driver's bench command and a short reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/probe_ab.py 0 2 > gpurun_out/probe_ab.log 2>&1; echo "probe_ab rc=$?"; cat gpurun_out/probe_ab.log
timeout 200 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; echo "latency rc=$?"; grep pair_c gpurun_out/probe_latency.log
if [ $rc -eq 0 ]; then
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$? wall=${SECONDS}s"; tail -c 600 gpurun_out/bench_ours.err
SECONDS=0
timeout 600 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_ours.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms_per_step", d["ms_per_step"], "us/step", d["roofline"]["step_kernel"]["us_per_scheduler_step"])
print("e2e", d["e2e"]["value"], "driver_ms", d["e2e"]["driver_ms_per_replay"], "retention", d["e2e"]["with_retention"]["value"], "no_zero", d["e2e"]["no_zero_page"].get("value"))
print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"].get("replays_s"), d["cpu_baseline"].get("prealloc_off_value"), d["cpu_baseline"].get("bookkeeping_only_value"))
print("roofline", d["roofline"]["frac"], "micro", {k: round(v, 1) for k, v in d["micro_alloc_free"].items() if isinstance(v, float)})
print("per_call", d["e2e_per_call_api"]["value"], "mla", {k: v.get("ours_vs_reference") for k, v in d["mla_pin"].items() if isinstance(v, dict)})
print("map_unmap", {k: (v.get("map_plus_unmap_vs_raw", v) if isinstance(v, dict) else v) for k, v in d["map_unmap"].items() if k.startswith("n")})
r = json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["ms_per_step"])
PY
fi
