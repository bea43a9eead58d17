#!/bin/bash
# round 2, session 11 (8 GPUs): multi-GPU TP tests, bench N=8 (mailbox), reference arm N=8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_tp_mailbox_gpu.py -m gpu -x -q > gpurun_out/tests_tp_8gpu.log 2>&1; echo "tp pytest rc=$?"; tail -4 gpurun_out/tests_tp_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench N=8 mailbox rc=$?"
tail -c 2500 gpurun_out/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "reference N=8 rc=$?"
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        return {"error": str(e)}
d = last("gpurun_out/bench_n8.json")
if "value" in d:
    print("N=8 value", d["value"], "e2e", d["e2e"]["value"], "retention", d["e2e"]["with_retention"]["value"], "tp", d["e2e"]["tp"])
    print(" strict", d["e2e"]["strict"])
    print(" scatter", d["kv_scatter"]["scatter"], d["kv_scatter"]["gather"])
    print(" map_unmap", {k: (v.get("map_plus_unmap_vs_raw", v) if isinstance(v, dict) else v) for k, v in d["map_unmap"].items()})
else:
    print(d)
r = last("gpurun_out/bench_ref_n8.json")
print("ref N=8", r.get("value"), r.get("cpu_baseline", {}).get("sample", "")[:200])
PY
