#!/bin/bash
# round 2, pass f (2 GPUs): the default bench line under torchrun (routed exchange), the exact fused
# exchange beside it, and the multi-rank CLI at full size against the oracle's digest
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err
tail -c 1500 gpurun_out/r02f_bench_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload indexer --exchange fused > gpurun_out/r02f_bench_${N}gpu_k15_fused_exact.json 2> gpurun_out/r02f_bench_${N}gpu_k15_fused_exact.err
timeout 900 $TR tools/cli_e2e_mr.py 1.0 15 > gpurun_out/r02f_cli_${N}gpu_k15.json 2> gpurun_out/r02f_cli_${N}gpu_k15.err
tail -c 800 gpurun_out/r02f_cli_${N}gpu_k15.err; cat gpurun_out/r02f_cli_${N}gpu_k15.json | tail -n 3
python - <<PY
import json
for f in ("bench_${N}gpu", "bench_${N}gpu_k15_fused_exact"):
    try:
        l = json.loads(open(f"gpurun_out/r02f_{f}.json").read().strip().splitlines()[-1])
        print(f, l["ms_per_step"], l["value"], l["config"].get("exact_redo_steps"), l["roofline"]["kernel_ms_by_class"], "e2e", l["e2e"]["ms_per_step"] if l.get("e2e") else None, l.get("parity_check"))
        for k in ("indexer_k17", "indexer_k19"):
            if k in l: print(" ", k, l[k]["ms_per_step"], l[k].get("parity_check"), l[k]["config"].get("exact_redo_steps"))
        for k, m in l.get("merger", {}).items(): print(" ", k, m["ms_per_step"], m["parity_check"], "e2e", m["e2e"]["ms_per_step"] if m.get("e2e") else None)
    except Exception as e:
        print(f, "failed", e)
PY
