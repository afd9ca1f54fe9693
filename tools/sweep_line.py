#!/usr/bin/env python3
"""Print one line (ms per step + per-class kernel times) from a bench.py log: tools/sweep_line.py LABEL LOG"""
import json
import sys

d = json.loads([l for l in open(sys.argv[2]).read().strip().splitlines() if l.startswith("{")][-1])
print(sys.argv[1], round(d["ms_per_step"], 3), "ms", d["roofline"].get("kernel_ms_by_class"))
