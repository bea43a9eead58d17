#!/bin/bash
# round 2, session 5: tests (background init, two-model leg), full bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 1500 python bench.py --steps 3 --warmup 3 --ref-replays 1 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_ours.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "with_retention", d["e2e"]["with_retention"]["value"])
    print("two_model", json.dumps(d["two_model"], indent=1))
    print("init", json.dumps(d["init_and_layout"], indent=1))
    print("per_call", d["e2e_per_call_api"]["value"])
except Exception as e:
    print("no bench line:", e)
PY
