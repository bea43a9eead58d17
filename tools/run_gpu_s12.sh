#!/bin/bash
# round 2, session 12 (4 GPUs): bench N=4 (mailbox) + reference arm N=4
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench N=4 mailbox rc=$?"
tail -c 1500 gpurun_out/bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29732 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err; echo "reference N=4 rc=$?"
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        return {"error": str(e)}
d = last("gpurun_out/bench_n4.json")
if "value" in d:
    print("N=4 value", d["value"], "e2e", d["e2e"]["value"], "retention", d["e2e"]["with_retention"]["value"], "tp", d["e2e"]["tp"])
    print(" scatter", d["kv_scatter"]["scatter"])
    print(" map_unmap", {k: (v.get("map_plus_unmap_vs_raw", v) if isinstance(v, dict) else v) for k, v in d["map_unmap"].items()})
else:
    print(d)
print("ref N=4", last("gpurun_out/bench_ref_n4.json").get("value"))
PY
