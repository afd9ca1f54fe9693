#!/bin/bash
# round 2, pass u (N GPUs): the committed state under torchrun -- default line (K=15 routed, K=17, K=19 at >= 4 GPUs, merger)
mkdir -p gpurun_out
N=${1:-8}
nproc; free -g | head -2 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02u_bench_${N}gpu.json 2> gpurun_out/r02u_bench_${N}gpu.err
tail -c 600 gpurun_out/r02u_bench_${N}gpu.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r02u_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("K15", round(l["ms_per_step"], 3), l["value"], l["config"].get("exact_redo_steps"), l["roofline"]["kernel_ms_by_class"], "e2e", l["e2e"], l.get("parity_check"))
print("   per_rank", l["config"].get("per_rank"))
for k in ("indexer_k17", "indexer_k19"):
    if k in l: print(" ", k, round(l[k]["ms_per_step"], 3), l[k].get("parity_check"), "e2e", l[k]["e2e"]["ms_per_step"] if l[k].get("e2e") else None, (l[k]["e2e"] or {}).get("d2h"))
for k, m in l.get("merger", {}).items(): print(" ", k, round(m["ms_per_step"], 4), m["parity_check"]["equal"], "e2e", m["e2e"]["ms_per_step"] if m.get("e2e") else None)
PY
