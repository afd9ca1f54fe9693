#!/bin/bash
# round 2, pass y (1 GPU): a change of the scan front end against the library built before it (A/B in one call)
mkdir -p gpurun_out
cp pykmer_b200/libpykmer_b200.so /tmp/lib_new.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_streams or chunked or records or sequence_sharded or routed or direct or partition or byte_windows" > gpurun_out/r02y_pytest_indexer.log 2>&1; tail -n 2 gpurun_out/r02y_pytest_indexer.log
for v in before new before new; do
  cp /tmp/lib_$v.so pykmer_b200/libpykmer_b200.so 2>/dev/null || cp build/variants/lib_$v.so pykmer_b200/libpykmer_b200.so
  for k in 15 17; do
    timeout 600 python bench.py --workload indexer --kmer $k --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02y_tmp.json 2> gpurun_out/r02y_tmp.err
    python -c "
import json; l=json.loads(open('gpurun_out/r02y_tmp.json').read().strip().splitlines()[-1]); c=l['roofline']['kernel_ms_by_class']; print('$v K=$k', round(l['ms_per_step'],3), 'pass1', c['scan_bucket_count'], 'pass2', c['scan_scatter'], l['parity_check']['equal'])"
  done
done | tee gpurun_out/r02y_encoder_ab.txt
cp /tmp/lib_new.so pykmer_b200/libpykmer_b200.so
