#!/bin/bash
# round 2, pass k (1 GPU): Gram kernel with one mbarrier arrival per producer warp
mkdir -p gpurun_out
SWEEP_NS=3,50,100,128,255 SWEEP_VARIANTS=tmem timeout 600 python tools/gram_sweep.py > gpurun_out/r02k_gram_sweep.txt 2>&1
grep -E "diag=0|diag=3|diag=4" gpurun_out/r02k_gram_sweep.txt
for n in 255 253 131 50; do DIAG_BITS=0 DIAG_SHORT=1 DIAG_N=$n DIAG_REPS=8 DIAG_LOGW=15,23 timeout 300 python tools/gram_diag.py; done 2>&1 | grep -v "rows \[\] cols \[\]" | tail -20 > gpurun_out/r02k_gram_diag.txt; echo "diag lines with mismatches: $(wc -l < gpurun_out/r02k_gram_diag.txt)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or tiled or merge or merger or pack" > gpurun_out/r02k_pytest_merger.log 2>&1; tail -n 2 gpurun_out/r02k_pytest_merger.log
timeout 900 python -m pytest tests/test_gpu_at_scale.py -m gpu -x -q -k "merger or f4" > gpurun_out/r02k_pytest_at_scale_merger.log 2>&1; tail -n 2 gpurun_out/r02k_pytest_at_scale_merger.log
for n in 50 255; do
  timeout 600 python bench.py --workload merger --samples $n --max-count $n --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_merger_n$n.json 2> gpurun_out/r02k_bench_merger_n$n.err
done
python - <<'PY'
import json
for f in ("merger_n50", "merger_n255"):
    try:
        l = json.loads(open(f"gpurun_out/r02k_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, l["ms_per_step"], l.get("parity_check", {}).get("equal"))
    except Exception as e:
        print(f, "failed", e)
PY
