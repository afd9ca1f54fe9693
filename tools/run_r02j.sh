#!/bin/bash
# round 2, pass j (2 GPUs): the final code -- routed exchange with the status words carrying num_kmers,
# per-rank kernel times in the line
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "routed or fused_exchange or cli_two_ranks" > gpurun_out/r02j_pytest_routed.log 2>&1; tail -n 3 gpurun_out/r02j_pytest_routed.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02j_bench_${N}gpu.json 2> gpurun_out/r02j_bench_${N}gpu.err
tail -c 800 gpurun_out/r02j_bench_${N}gpu.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r02j_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("K15", l["ms_per_step"], l["value"], l["config"].get("exact_redo_steps"), l["roofline"]["kernel_ms_by_class"], "e2e", l["e2e"]["ms_per_step"] if l.get("e2e") else None, l.get("parity_check"))
print("   per_rank", l["config"].get("per_rank"))
for k in ("indexer_k17", "indexer_k19"):
    if k in l: print(" ", k, l[k]["ms_per_step"], l[k].get("parity_check"), l[k]["config"].get("exact_redo_steps"))
for k, m in l.get("merger", {}).items(): print(" ", k, m["ms_per_step"], m["parity_check"]["equal"], "e2e", m["e2e"]["ms_per_step"] if m.get("e2e") else None)
PY
