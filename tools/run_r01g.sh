# Round-1: tiled masks + FP4 Gram as the merger's default -- the tests that touch it, then bench lines.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests -q -m gpu \
    -k "two_ranks or gram or merge or merger or cli or distance or stats or synth_table or pack or tiled" \
    > gpurun_out/r01g_pytest_subset.log 2>&1
tail -12 gpurun_out/r01g_pytest_subset.log
for n in 50 255; do
  timeout 150 python bench.py --workload merger --samples $n --max-count $([ $n = 50 ] && echo 50 || echo 255) \
      --no-cpu-baseline --no-e2e \
      > gpurun_out/r01g_bench_merger_n${n}.json 2> gpurun_out/r01g_bench_merger_n${n}.err
  echo "merger n=$n: $(head -c 330 gpurun_out/r01g_bench_merger_n${n}.json)"; tail -2 gpurun_out/r01g_bench_merger_n${n}.err
done
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01g_smoke.log 2>&1; tail -1 gpurun_out/r01g_smoke.log
