#!/bin/bash
# round 2, pass q (1 GPU): k_scan_scatter instruction diet
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_streams or finalize_to_host or sequence_sharded or routed" > gpurun_out/r02q_pytest_indexer.log 2>&1; tail -n 2 gpurun_out/r02q_pytest_indexer.log
for k in 15 17; do
timeout 600 python bench.py --workload indexer --kmer $k --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02q_bench_k$k.json 2> gpurun_out/r02q_bench_k$k.err
done
python - <<'PY'
import json
for k in (15, 17):
    try:
        l = json.loads(open(f"gpurun_out/r02q_bench_k{k}.json").read().strip().splitlines()[-1])
        print(k, round(l["ms_per_step"], 3), l["roofline"]["kernel_ms_by_class"], (l.get("parity_check") or {}).get("equal"))
    except Exception as e:
        print(k, "failed", e)
PY
