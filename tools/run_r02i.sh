#!/bin/bash
# round 2, pass i (N GPUs): the default bench line under torchrun (K=15 routed, K=17, K=19, merger) and
# the multi-rank CLI at full size against the oracle's digest
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02i_bench_${N}gpu.json 2> gpurun_out/r02i_bench_${N}gpu.err
tail -c 1200 gpurun_out/r02i_bench_${N}gpu.err
[ "$N" = "8" ] && timeout 600 $TR tools/cli_e2e_mr.py 1.0 15 > gpurun_out/r02i_cli_${N}gpu_k15.json 2> gpurun_out/r02i_cli_${N}gpu_k15.err
tail -c 600 gpurun_out/r02i_cli_${N}gpu_k15.err; tail -n 2 gpurun_out/r02i_cli_${N}gpu_k15.json
python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r02i_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("K15", l["ms_per_step"], l["value"], l["config"].get("exact_redo_steps"), l["roofline"]["kernel_ms_by_class"], "e2e", l["e2e"]["ms_per_step"] if l.get("e2e") else None, l.get("parity_check"))
    for k in ("indexer_k17", "indexer_k19"):
        if k in l: print(" ", k, l[k]["ms_per_step"], l[k].get("parity_check"), l[k]["config"].get("exact_redo_steps"), l[k]["roofline"]["kernel_ms_by_class"], "e2e", l[k]["e2e"]["ms_per_step"] if l[k].get("e2e") else None)
    for k, m in l.get("merger", {}).items(): print(" ", k, m["ms_per_step"], m["parity_check"], "e2e", m["e2e"]["ms_per_step"] if m.get("e2e") else None)
except Exception as e:
    print("failed", e)
PY
