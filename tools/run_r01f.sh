# Round-1 last GPU call: the tests of what changed since r01e (merger driver, Gram kernels, the CLIs
# as multi-rank jobs, Header helpers), the opt-in FP4 Gram experiment, and i8 / f4 merger bench lines
# side by side on the same box.  Every step under its own timeout; nothing here is a bench value
# taken under a profiler.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PYKMER_B200_TEST_F4=1 timeout 330 python -m pytest tests -q -m gpu \
    -k "two_ranks or gram or merge or merger or cli or distance or stats or synth_table or pack" \
    > gpurun_out/r01f_pytest_subset.log 2>&1
tail -15 gpurun_out/r01f_pytest_subset.log
for algo in i8 f4; do
  for n in 50 255; do
    PYKMER_B200_GRAM=$algo timeout 150 python bench.py --workload merger --samples $n \
        --max-count $([ $n = 50 ] && echo 50 || echo 255) --no-e2e --no-cpu-baseline \
        > gpurun_out/r01f_bench_merger_n${n}_${algo}.json 2> gpurun_out/r01f_bench_merger_n${n}_${algo}.err
    echo "merger n=$n $algo: $(head -c 260 gpurun_out/r01f_bench_merger_n${n}_${algo}.json)"
  done
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01f_smoke.log 2>&1; tail -1 gpurun_out/r01f_smoke.log
