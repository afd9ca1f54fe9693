#!/bin/bash
# round 2, pass p (1 GPU): ncu captures with source of k_scan_scatter (where do 115 thread-instructions per base go)
# and of k_table_pack (the new kernel of the packed table transfer)
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py 0.25 15 0 2 > gpurun_out/r02p_plain_step.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_scan_scatter' -s 2 -c 1 -f -o gpurun_out/r02p_prof_scan_scatter python tools/profile_step.py 0.25 15 0 2 > gpurun_out/r02p_ncu_a.log 2>&1
PROBE_LOG2=26 timeout 300 python tools/pack_probe.py > gpurun_out/r02p_pack_probe.txt 2>&1
PROBE_LOG2=26 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_table_pack -c 1 -s 2 -f -o gpurun_out/r02p_prof_table_pack python tools/pack_probe.py > gpurun_out/r02p_ncu_b.log 2>&1
tail -n 2 gpurun_out/r02p_ncu_?.log; cat gpurun_out/r02p_pack_probe.txt | head -12
