#!/bin/bash
# round 2, pass d: FP4 Gram kernel v4 with the raw-slot release fixed (LDS + fence before mbarrier.arrive)
mkdir -p gpurun_out
SWEEP_NS=3,50,100,128,255 SWEEP_VARIANTS=tmem timeout 600 python tools/gram_sweep.py > gpurun_out/r02d_gram_sweep.txt 2>&1
PYKMER_B200_GRAM_DIAG=8 timeout 300 python bench.py --workload merger --samples 255 --max-count 255 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > gpurun_out/r02d_bench_merger_n255_unfused.json 2> gpurun_out/r02d_bench_merger_n255_unfused.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or tiled or merge or merger or pack" > gpurun_out/r02d_pytest_merger.log 2>&1
timeout 900 python -m pytest tests/test_gpu_at_scale.py -m gpu -x -q -k "merger or f4" > gpurun_out/r02d_pytest_at_scale_merger.log 2>&1
timeout 600 python bench.py --workload merger --samples 255 --max-count 255 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02d_bench_merger_n255.json 2> gpurun_out/r02d_bench_merger_n255.err
timeout 600 python bench.py --workload merger --samples 50 --max-count 50 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02d_bench_merger_n50.json 2> gpurun_out/r02d_bench_merger_n50.err
tail -n 3 gpurun_out/r02d_pytest_merger.log gpurun_out/r02d_pytest_at_scale_merger.log
cat gpurun_out/r02d_gram_sweep.txt
python - <<'PY'
import json
for f in ("merger_n255_unfused", "merger_n255", "merger_n50"):
    try:
        l = json.loads(open(f"gpurun_out/r02d_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, l["ms_per_step"], l.get("parity_check"))
    except Exception as e:
        print(f, "failed", e)
PY
DIAG_BITS=0,8 DIAG_SHORT=1 DIAG_N=255 DIAG_REPS=6 timeout 300 python tools/gram_diag.py 2>&1 | grep -v "rows \[\] cols \[\]" | tail -20 > gpurun_out/r02d_gram_diag.txt; echo "diag lines with mismatches: $(wc -l < gpurun_out/r02d_gram_diag.txt)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "routed or fused_exchange or sequence_sharded" > gpurun_out/r02d_pytest_routed.log 2>&1; tail -n 5 gpurun_out/r02d_pytest_routed.log
