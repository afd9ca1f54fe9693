#!/bin/bash
# round 2, pass o (2 GPUs): the default line under torchrun with the packed table transfer on every rank
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
nproc; free -g | head -2 | tail -1
for packed in 1 0; do
PYKMER_B200_PACKED_D2H=$packed timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-reference-python > gpurun_out/r02o_bench_${N}gpu_packed$packed.json 2> gpurun_out/r02o_bench_${N}gpu_packed$packed.err
tail -c 400 gpurun_out/r02o_bench_${N}gpu_packed$packed.err
done
python - <<PY
import json
for packed in (1, 0):
    l = json.loads(open(f"gpurun_out/r02o_bench_${N}gpu_packed{packed}.json").read().strip().splitlines()[-1])
    print("packed", packed, "K15", round(l["ms_per_step"], 3), "e2e", l["e2e"], l.get("parity_check"))
    for k in ("indexer_k17", "indexer_k19"):
        if k in l: print("   ", k, round(l[k]["ms_per_step"], 3), l[k].get("parity_check"), "e2e", l[k]["e2e"]["ms_per_step"] if l[k].get("e2e") else None, l[k]["e2e"].get("d2h") if l[k].get("e2e") else None)
    for k, m in l.get("merger", {}).items(): print("   ", k, round(m["ms_per_step"], 3), m["parity_check"]["equal"])
PY
