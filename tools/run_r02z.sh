#!/bin/bash
# round 2, pass z (1 GPU): final state -- full GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest_gpu_full.log 2>&1; tail -n 3 gpurun_out/r02z_pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02z_smoke.log 2>&1; tail -n 1 gpurun_out/r02z_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err; tail -c 300 gpurun_out/r02z_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference_arm.json 2> gpurun_out/r02z_bench_reference_arm.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02z_bench_default.json").read().strip().splitlines()[-1])
print("K15", round(l["ms_per_step"], 3), l["roofline"]["kernel_ms_by_class"], "step_frac", round(l["roofline"]["step_frac"], 3), "e2e", round(l["e2e"]["ms_per_step"], 2), l["e2e"]["d2h"]["packed_windows"], l["e2e"].get("host_table_sha256_equals_oracle"), l["parity_check"]["equal"], "cpu", l.get("cpu_baseline", {}).get("value"))
s = l["indexer_k17"]; print("K17", round(s["ms_per_step"], 3), s["parity_check"]["equal"], "e2e", round(s["e2e"]["ms_per_step"], 1), "step_frac", round(s["roofline"]["step_frac"], 3))
for k, m in l["merger"].items(): print(k, round(m["ms_per_step"], 3), m["parity_check"]["equal"], "e2e", m["e2e"]["ms_per_step"] if m.get("e2e") else None, m["roofline"]["frac"])
r = json.loads(open("gpurun_out/r02z_bench_reference_arm.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["config"]["workload"] == l["config"]["workload"])
PY
