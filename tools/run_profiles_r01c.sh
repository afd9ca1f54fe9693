set -x
python bench.py > gpurun_out/r01c_bench_default.log 2> gpurun_out/r01c_bench_default.err; tail -c 600 gpurun_out/r01c_bench_default.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01c_bench_reference.log 2>&1; tail -c 400 gpurun_out/r01c_bench_reference.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01c_launches_k15.csv python tools/profile_step.py 1.0 15 0 2 > gpurun_out/r01c_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_window_count$' -s 8 -c 2 -f -o gpurun_out/r01c_prof_window python tools/profile_step.py 1.0 15 0 1 > gpurun_out/r01c_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_scatter -c 1 -f -o gpurun_out/r01c_prof_scatter python tools/profile_step.py 1.0 15 0 1 > gpurun_out/r01c_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_window_count8 -s 8 -c 2 -f -o gpurun_out/r01c_prof_window8 python tools/profile_step.py 1.0 17 0 1 > gpurun_out/r01c_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_count_direct -c 1 -f -o gpurun_out/r01c_prof_direct python tools/profile_step.py 1.0 19 0 1 96:128 > gpurun_out/r01c_ncu_d.log 2>&1
python bench.py --kmer 17 > gpurun_out/r01c_bench_k17.log 2> gpurun_out/r01c_bench_k17.err; tail -c 300 gpurun_out/r01c_bench_k17.log
python bench.py --workload merger --samples 50 --max-count 50 > gpurun_out/r01c_merger_n50.log 2> gpurun_out/r01c_merger_n50.err; tail -c 300 gpurun_out/r01c_merger_n50.log
python bench.py --workload merger --samples 255 --max-count 255 > gpurun_out/r01c_merger_n255.log 2> gpurun_out/r01c_merger_n255.err; tail -c 300 gpurun_out/r01c_merger_n255.log
ls -la gpurun_out/*.ncu-rep
