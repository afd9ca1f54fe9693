set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r01e_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01e_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01e_smoke.log 2>&1; tail -1 gpurun_out/r01e_smoke.log
python bench.py > gpurun_out/r01e_bench_default.log 2> gpurun_out/r01e_bench_default.err; tail -c 300 gpurun_out/r01e_bench_default.log
python bench.py --kmer 17 > gpurun_out/r01e_bench_k17.log 2> gpurun_out/r01e_bench_k17.err; tail -c 300 gpurun_out/r01e_bench_k17.log
python bench.py --kmer 11 --scale 0.0128 > gpurun_out/r01e_bench_config1_k11.log 2>&1; tail -c 300 gpurun_out/r01e_bench_config1_k11.log
