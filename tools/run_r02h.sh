#!/bin/bash
# round 2, pass h (1 GPU): expansion shifts on the FMA pipe (Gram sweep), smoke, the new GPU tests, and
# k_window_count in its PRODUCTION cache state: ncu application replay (no save / restore of the
# 64 MiB counter window between passes), a handful of metrics
mkdir -p gpurun_out
SWEEP_NS=3,50,100,128,255 SWEEP_VARIANTS=tmem timeout 600 python tools/gram_sweep.py > gpurun_out/r02h_gram_sweep.txt 2>&1
cat gpurun_out/r02h_gram_sweep.txt | grep -E "diag=0|diag=3|diag=4"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02h_smoke.log 2>&1; tail -n 2 gpurun_out/r02h_smoke.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or tiled or merge or merger or pack or cli_two_ranks or routed" > gpurun_out/r02h_pytest_subset.log 2>&1; tail -n 4 gpurun_out/r02h_pytest_subset.log
timeout 900 python -m pytest tests/test_gpu_at_scale.py -m gpu -x -q -k "merger or f4" > gpurun_out/r02h_pytest_at_scale_merger.log 2>&1; tail -n 2 gpurun_out/r02h_pytest_at_scale_merger.log
for n in 50 255; do
  timeout 600 python bench.py --workload merger --samples $n --max-count $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_merger_n$n.json 2> gpurun_out/r02h_bench_merger_n$n.err
done
python - <<'PY'
import json
for f in ("merger_n50", "merger_n255"):
    try:
        l = json.loads(open(f"gpurun_out/r02h_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, l["ms_per_step"], l.get("parity_check", {}).get("equal"), "e2e", l["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python tools/profile_step.py 1.0 15 0 3 > gpurun_out/r02h_plain_step.log 2>&1 &&
ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k 'regex:^k_window_count$' -s 134 -c 4 --csv --log-file gpurun_out/r02h_ncu_window_count_app_replay.csv python tools/profile_step.py 1.0 15 0 3 > gpurun_out/r02h_ncu_a.log 2>&1
tail -n 3 gpurun_out/r02h_ncu_a.log; head -c 1500 gpurun_out/r02h_ncu_window_count_app_replay.csv
