#!/bin/bash
# A/B in one call on one box: the library in the tree against build/variants/lib_before.so (K=15 / K=17 device steps)
mkdir -p gpurun_out
cp pykmer_b200/libpykmer_b200.so /tmp/lib_new.so; cp build/variants/lib_before.so /tmp/lib_before.so
for v in before new before new; do
  cp /tmp/lib_$v.so pykmer_b200/libpykmer_b200.so
  for k in ${AB_KS:-15}; do
    timeout 600 python bench.py --workload indexer --kmer $k --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
    python -c "
import json; l=json.loads(open('gpurun_out/ab_tmp.json').read().strip().splitlines()[-1]); c=l['roofline']['kernel_ms_by_class']; print('$v K=$k', round(l['ms_per_step'],3), 'pass1', c['scan_bucket_count'], 'pass2', c['scan_scatter'], l['parity_check']['equal'])"
  done
done | tee gpurun_out/ab_result.txt
cp /tmp/lib_new.so pykmer_b200/libpykmer_b200.so
