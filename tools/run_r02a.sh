#!/bin/bash
# round 2, pass a: the new FP4 Gram kernel (A from TMEM, dual slab, block pairs), the at-scale parity
# tests, the new default bench line.  Run on the GPU box: gpurun -- bash tools/run_r02a.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
free -g >> gpurun_out/r02a_gpu.txt; nproc >> gpurun_out/r02a_gpu.txt
timeout 120 ./tools/microbench3 > gpurun_out/r02a_microbench3.txt 2>&1
SWEEP_NS=3,50,100,255 timeout 600 python tools/gram_sweep.py > gpurun_out/r02a_gram_sweep.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or tiled or merge or merger or pack" > gpurun_out/r02a_pytest_merger.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_at_scale.py -m gpu -q --durations=20 > gpurun_out/r02a_pytest_at_scale.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_pytest_merger.log gpurun_out/r02a_pytest_at_scale.log
cat gpurun_out/r02a_gram_sweep.txt
