#!/bin/bash
# round 2, pass l (1 GPU): packed device-to-host transfer of the table (k_table_pack + host unpack team),
# Gram kernel arrivals per warp at 129..256 rows only
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "finalize_to_host or table_pack or gram or tiled or merge" > gpurun_out/r02l_pytest_subset.log 2>&1; tail -n 3 gpurun_out/r02l_pytest_subset.log
timeout 900 python -m pytest tests/test_gpu_at_scale.py -m gpu -x -q -k "config2 or f4 or merger" > gpurun_out/r02l_pytest_at_scale.log 2>&1; tail -n 3 gpurun_out/r02l_pytest_at_scale.log
for cfg in "packed:1:" "raw:0:" "packed_t7:1:7" "packed_t11:1:11" "packed_t15:1:15"; do
  IFS=: read name packed threads <<< "$cfg"
  PYKMER_B200_PACKED_D2H=$packed PYKMER_B200_UNPACK_THREADS=$threads timeout 600 python bench.py --workload indexer --kmer 15 --steps 5 --warmup 3 --no-cpu-baseline \
     > gpurun_out/r02l_bench_k15_$name.json 2> gpurun_out/r02l_bench_k15_$name.err
done
PYKMER_B200_PACKED_D2H=1 timeout 600 python bench.py --workload indexer --kmer 17 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_k17_packed.json 2> gpurun_out/r02l_bench_k17_packed.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02l_bench_k1*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step", round(l["ms_per_step"], 3), "e2e", l["e2e"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
SWEEP_NS=50,128,255 SWEEP_VARIANTS=tmem SWEEP_DIAGS=0 timeout 300 python tools/gram_sweep.py 2>&1 | grep "diag=0"
