#!/bin/bash
# round 2, pass e: FP4 Gram kernel v4, TMA L2 promotion 128 B (odd sample counts), raw-slot release fenced
mkdir -p gpurun_out
SWEEP_NS=3,50,100,128,255 SWEEP_VARIANTS=tmem timeout 600 python tools/gram_sweep.py > gpurun_out/r02e_gram_sweep.txt 2>&1
PYKMER_B200_GRAM_DIAG=8 timeout 300 python bench.py --workload merger --samples 255 --max-count 255 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > gpurun_out/r02e_bench_merger_n255_unfused.json 2> gpurun_out/r02e_bench_merger_n255_unfused.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or tiled or merge or merger or pack" > gpurun_out/r02e_pytest_merger.log 2>&1
timeout 900 python -m pytest tests/test_gpu_at_scale.py -m gpu -x -q -k "merger or f4" > gpurun_out/r02e_pytest_at_scale_merger.log 2>&1
timeout 600 python bench.py --workload merger --samples 255 --max-count 255 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_bench_merger_n255.json 2> gpurun_out/r02e_bench_merger_n255.err
timeout 600 python bench.py --workload merger --samples 50 --max-count 50 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_bench_merger_n50.json 2> gpurun_out/r02e_bench_merger_n50.err
tail -n 3 gpurun_out/r02e_pytest_merger.log gpurun_out/r02e_pytest_at_scale_merger.log
cat gpurun_out/r02e_gram_sweep.txt
python - <<'PY'
import json
for f in ("merger_n255_unfused", "merger_n255", "merger_n50"):
    try:
        l = json.loads(open(f"gpurun_out/r02e_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, l["ms_per_step"], l.get("parity_check"))
    except Exception as e:
        print(f, "failed", e)
PY
for n in 255 253 131; do DIAG_BITS=0,8 DIAG_SHORT=1 DIAG_N=$n DIAG_REPS=8 DIAG_LOGW=15,23 timeout 300 python tools/gram_diag.py; done 2>&1 | grep -v "rows \[\] cols \[\]" | tail -20 > gpurun_out/r02e_gram_diag.txt; echo "diag lines with mismatches: $(wc -l < gpurun_out/r02e_gram_diag.txt)"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu_full.log 2>&1; tail -n 5 gpurun_out/r02e_pytest_gpu_full.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 600 gpurun_out/r02e_bench.err
