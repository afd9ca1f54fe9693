#!/bin/bash
# round 2, pass g (1 GPU): NP=128 Gram kernels with four producer groups; ncu evidence -- the launch
# list of an indexer step, k_window_count / k_window_commit in their PRODUCTION cache state
# (--cache-control none, windows of the third step), the new k_gram_f4 at N=255 and N=50
mkdir -p gpurun_out
SWEEP_NS=3,50,100,128,255 SWEEP_VARIANTS=tmem timeout 600 python tools/gram_sweep.py > gpurun_out/r02g_gram_sweep.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or tiled or merge or merger or pack" > gpurun_out/r02g_pytest_merger.log 2>&1
timeout 900 python -m pytest tests/test_gpu_at_scale.py -m gpu -x -q -k "merger or f4" > gpurun_out/r02g_pytest_at_scale_merger.log 2>&1
for n in 50 255; do
  timeout 600 python bench.py --workload merger --samples $n --max-count $n --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02g_bench_merger_n$n.json 2> gpurun_out/r02g_bench_merger_n$n.err
done
tail -n 3 gpurun_out/r02g_pytest_merger.log gpurun_out/r02g_pytest_at_scale_merger.log
cat gpurun_out/r02g_gram_sweep.txt
python - <<'PY'
import json
for f in ("merger_n50", "merger_n255"):
    try:
        l = json.loads(open(f"gpurun_out/r02g_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, l["ms_per_step"], l.get("parity_check", {}).get("equal"))
    except Exception as e:
        print(f, "failed", e)
PY
# ---- ncu (each only after the identical command has exited 0 without it)
python tools/profile_step.py 1.0 15 0 3 > gpurun_out/r02g_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_launches_k15.csv python tools/profile_step.py 1.0 15 0 3 > gpurun_out/r02g_ncu_list.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k 'regex:^k_window_count$' -s 134 -c 3 -f -o gpurun_out/r02g_prof_window_count_warm python tools/profile_step.py 1.0 15 0 3 > gpurun_out/r02g_ncu_a.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k 'regex:k_window_commit' -s 134 -c 2 -f -o gpurun_out/r02g_prof_window_commit_warm python tools/profile_step.py 1.0 15 0 3 > gpurun_out/r02g_ncu_b.log 2>&1
python tools/profile_merger.py 255 13 > gpurun_out/r02g_plain_merger255.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_gram_f4' -s 2 -c 1 -f -o gpurun_out/r02g_prof_gram_n255_k13 python tools/profile_merger.py 255 13 > gpurun_out/r02g_ncu_c.log 2>&1
python tools/profile_merger.py 50 13 > gpurun_out/r02g_plain_merger50.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_gram_f4' -s 2 -c 1 -f -o gpurun_out/r02g_prof_gram_n50_k13 python tools/profile_merger.py 50 13 > gpurun_out/r02g_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/r02g_ncu_?.log
