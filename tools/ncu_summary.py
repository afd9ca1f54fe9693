#!/usr/bin/env python3
"""Condense .ncu-rep files (read here, without a GPU) into the text tables kept under profiles/.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "l2_atomic_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {rep}: no data")
            continue
        hdr, units = rows[0], rows[1]
        print(f"# {rep}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"kernel: {name[:110]}")
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"    {label:14s} {r[i]:>18s} {units[i]}")
            stalls = []
            for i, h in enumerate(hdr):
                if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("    top stalls     " + ", ".join(f"{n}={int(v)}" for v, n in stalls[:6]))
        print()


if __name__ == "__main__":
    main()
