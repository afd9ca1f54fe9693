#!/usr/bin/env python3
"""bench.py -- every metric BASELINE.json names, on B200, in ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                             # CPU arm (oracle port)
    python bench.py --workload indexer|merger ...                    # one workload only (development)

The top level of the line is the headline: indexer bp/s at K=15 on the synthetic tomato-sized
782,520,033 bp multi-FASTA stream (BASELINE.json configs[1]).  The same line carries, as
sub-objects with the same keys (value, ms_per_step, roofline, e2e, cpu_baseline, parity_check):

    "indexer_k17"  configs[3], the 16 GiB table
    "indexer_k19"  configs[4], the 256 GiB table sharded over >= 4 GPUs
    "merger"       {"n50": configs[2] (--max-count=50), "n255": configs[3]}: the Gram stage in GB/s of
                   presence bitmask, roofline on ISSUED tensor operations and on the HBM floor

One "step" = one whole pass of the hot path over the workload: for the indexer, scan + count the
782.5 Mbp stream into a fresh table and compute hist / vals_* (and, for e2e, move the stream in from
pinned host memory and the table back out); for the merger, one Gram contraction over all samples.
N > 1 (torchrun): K <= 17 shards the SEQUENCE for the scan (each rank scans 1/N of the stream and
stores its k-mer entries straight into the buffer of the rank that owns their table window, over
NVLink) and the K-MER AXIS for the count; K = 19 replicates the scan and shards the k-mer axis in
balanced ranges; the merger shards the k-mer axis and all-reduces the N x N partial matrices.
Total work is fixed, so scaling is "strong".

Prints ONE JSON line (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = {"indexer": ("indexer_bp_per_s_K{K}", "bp/s"), "merger": ("merger_bitmask_GB_per_s", "GB/s")}
L2_NOTE = "inputs exceed L2: 0.78 GB stream + 1 GiB table per step vs 126 MB L2"


def workload_name(kind: str, K: int, bp: int = 0, N: int = 0, max_count: int = 0) -> str:
    """config.workload -- the SAME string in this repo's arm and in the reference arm."""
    if kind == "indexer":
        return (f"indexer K={K}, synthetic tomato-sized multi-FASTA stream ({bp} bp, 13 records), "
                f"4^{K}-byte table")
    return (f"merger K={K}, N={N} synthetic samples, --max-count={max_count}; Gram stage over "
            f"{N * 4 ** K / 8 / 1e9:.2f} GB of presence bitmask")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "indexer", "merger"],
                    help="'all' (default): the headline indexer K=15 line carrying indexer_k17 / indexer_k19 / "
                         "merger sub-objects; 'indexer' / 'merger': that workload alone (--kmer, --samples)")
    ap.add_argument("--kmer", type=int, default=15)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the 782.5 Mbp genome")
    ap.add_argument("--samples", type=int, default=50, help="merger: number of samples")
    ap.add_argument("--max-count", type=int, default=50, help="merger: --max-count")
    ap.add_argument("--mode", type=int, default=0, help="indexer counting mode (0 auto)")
    ap.add_argument("--exchange", default="routed", choices=["routed", "fused", "nccl"],
                    help="sequence sharding: 'routed' = ONE scan pass stores into fixed regions of the owners' "
                         "buffers over NVLink (CUDA IPC peer mappings), sized once by the planning scan, no host "
                         "round trip inside a step; 'fused' = exact two-pass form (counts all-gathered through the "
                         "host every step), 'nccl' = bucket locally, then torch all_to_all_single")
    ap.add_argument("--shard", default="auto", choices=["auto", "sequence", "kmer"],
                    help="N > 1 indexer: 'sequence' = each rank scans 1/N of the stream and the k-mer entries "
                         "go to the window owners; 'kmer' = every rank scans everything, keeps its k-mer range")
    ap.add_argument("--emulate-shard", default=None, metavar="R/N | LO:HI",
                    help="development aid: on ONE GPU, run the k-mer-range shard that rank R of N would own "
                         "(not a bench line: one rank's share of a multi-GPU job)")
    ap.add_argument("--sub-steps", type=int, default=5, help="timed steps of the sub-object workloads (<= --steps)")
    ap.add_argument("--cpu-sample-mbp", type=float, default=1e9,
                    help="CPU leg: bases of the stream the port is timed on (default: all of it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-reference-python", action="store_true",
                    help="skip timing the real Python reference (baseline/_ref) on config 1")
    ap.add_argument("--no-parity-check", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------ inputs

def load_stream(scale: float, rank: int, world: int):
    """The cleaned 782.5 Mbp stream (+ record table).  Generated once per box (seeded,
    pykmer_b200/synth.py) and cached under the temp dir so that N ranks share it."""
    from pykmer_b200 import synth
    tag = f"pykmer_b200_syn782M_{synth.SYN782M_SEED:x}_{scale:.6f}"
    path = os.path.join(tempfile.gettempdir(), tag + ".npz")
    if rank == 0 and not os.path.exists(path):
        recs = synth.syn782m_records(scale=scale)
        stream, starts, lengths, names = synth.records_to_stream(recs)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, stream=stream, starts=starts, lengths=np.asarray(lengths, dtype=np.int64))
        os.replace(tmp, path)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    z = np.load(path)
    return z["stream"], z["starts"], z["lengths"].tolist()


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (every
    ~2 ms -- the timed regions here last tens of milliseconds, too short for `nvidia-smi -lms`),
    `nvidia-smi` as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.thread, self.stop_flag, self.sm, self.mx, self.reasons = None, False, [], [], set()
        self.timed = False                                   # samples are kept only while this is set
        self.near, self.near_reasons = [], set()             # ... and, apart, those of the warm-up steps

    def _poll(self, nv, handle):
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
            except Exception:
                break
            sm, reasons = (self.sm, self.reasons) if self.timed else (self.near, self.near_reasons)
            sm.append(float(mhz))
            for name, bit in names:
                if bits & bit:
                    reasons.add(name)
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: match the CUDA device by PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.idx), "pci_bus_id") else None
            handle = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    h = nv.nvmlDeviceGetHandleByIndex(i)
                    if int(nv.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        handle = h
                        break
            if handle is None:
                handle = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))]
            self.thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.timed = True
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.sm and self.near:
                # a timed region of a few milliseconds can fall between two polls: report the samples
                # of the warm-up steps of the same kernel that ran immediately before it, and say so
                tail = self.near[-16:]
                return {"sm_mhz": statistics.median(tail), "sm_max_mhz": self.mx[0],
                        "reasons": sorted(self.near_reasons), "samples": len(tail),
                        "how": "NVML, polled every ~2 ms; the timed region was shorter than one poll, these "
                               "are the samples of the warm-up steps immediately before it"}
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.mx[0] if self.mx else None, "reasons": ["no samples"]}
            return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.mx[0], "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "how": "NVML, polled every ~2 ms inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "how": "nvidia-smi -lms 100"}



def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/), or None."""
    try:
        import glob
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
        return int(json.load(open(files[-1]))[kernel]) if files else None
    except Exception:
        return None


def measured_peaks() -> dict:
    """Roofline denominators: MEASURED_PEAKS.json (driver-written), else the profiling recipe's fallback."""
    out = {"hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md 6.65 TB/s)",
           "bf16": 1590.0, "bf16_src": "fallback (B200_PROFILING.md bf16 1.59 PFLOP/s)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            out["hbm"], out["hbm_src"] = float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            out["bf16"], out["bf16_src"] = float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return out


def measured_peak():
    p = measured_peaks()
    return p["hbm"], p["hbm_src"]


# ------------------------------------------------------------------------------ CPU arm

def cpu_indexer_sample(stream: np.ndarray, K: int, sample_mbp: float, table=None):
    """The oracle port (oracle/kmer_oracle.c, threaded rolling form) with the full 4^K table on all
    host cores.  By default on the WHOLE stream (no extrapolation); with a smaller sample_mbp the
    per-base part (scan + count) is timed on a prefix and scaled, the per-table part (zero +
    hist/vals_* pass) is timed in full: t_job = t_scan * L/n + t_table."""
    from oracle import oracle
    n = int(min(stream.size, sample_mbp * 1e6))
    sample = stream if n == stream.size else np.ascontiguousarray(stream[:n])
    cores = oracle.max_threads()
    if table is None:
        table = np.zeros(4 ** K, dtype=np.uint8)
    t0 = time.perf_counter()
    table[:] = 0
    t1 = time.perf_counter()
    oracle.index_stream(sample, K, method="mt", threads=cores, table=table)
    t2 = time.perf_counter()
    oracle.table_stats(table, threads=cores)
    t3 = time.perf_counter()
    t_job = (t2 - t1) * stream.size / n + (t1 - t0) + (t3 - t2)
    if n == stream.size:
        what = (f"the whole {stream.size / 1e6:.1f} Mbp stream: zero the 4^{K}-byte table, scan + count, "
                f"hist / vals_* pass ({t_job:.2f} s on {cores} threads)")
    else:
        what = (f"scan+count timed on the first {n / 1e6:.1f} Mbp of the same stream ({t2 - t1:.2f} s, "
                f"{cores} threads) and scaled to {stream.size / 1e6:.1f} Mbp; zero + stats pass over the "
                f"4^{K}-byte table timed in full ({(t1 - t0) + (t3 - t2):.2f} s)")
    return stream.size / t_job, cores, what, (t3 - t0)


_MERGER_SAMPLE = {}


def cpu_merger_sample(K, N, max_count, repeats=24):
    """Pair loop of the reference (C port, all host cores) on a bounded slice of the k-mer axis
    and a subset of the samples, extrapolated to all N(N-1)/2 pairs over 4^K positions."""
    from oracle import oracle
    from pykmer_b200 import synth
    cores = oracle.max_threads()
    T = 4 ** K
    n = min(T, 1 << 22)
    Ns = min(N, 12)
    key = (K, Ns, n)
    if key not in _MERGER_SAMPLE:
        _MERGER_SAMPLE[key] = np.stack([synth.synth_table_slice(s, 0, n) for s in range(Ns)])
    tables = _MERGER_SAMPLE[key]
    oracle.merge_matrix(tables[:2], 1, max_count, threads=cores)      # warm the thread pool / pages
    t0 = time.perf_counter()
    for _ in range(repeats):
        oracle.merge_matrix(tables, 1, max_count, threads=cores)
    dt = (time.perf_counter() - t0) / repeats
    pair_positions = Ns * (Ns - 1) / 2 * n / dt
    full_pairs = N * (N - 1) / 2
    t_full = full_pairs * T / pair_positions
    what = (f"{Ns} samples x first {n} k-mers, all {Ns * (Ns - 1) // 2} pairs, {repeats} repeats of {dt:.3f} s on {cores} "
            f"threads; extrapolated to {int(full_pairs)} pairs x 4^{K} positions ({t_full:.0f} s)")
    return N * T / 8 / t_full / 1e9, cores, what, t_full


def reference_python():
    """The REAL reference (sauloal/pykmer's own indexer.py, unmodified, CPython, one core -- it is
    single-threaded by construction) timed on BASELINE config 1: the 10 Mbp bgzip multi-FASTA at K=11.
    It runs from baseline/_ref (a verbatim, git-ignored copy made by __graft_entry__.build() in the
    build container) through oracle/ref_shim.py (three non-arithmetic shims), in a
    subprocess; its .kin must hash to the committed golden.  Cached per box (both arms report it)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not all(os.path.exists(os.path.join(ref, f)) for f in ("indexer.py", "tools.py")):
        return {"unavailable": "baseline/_ref holds no copy of the reference (run __graft_entry__.build() "
                               "where /root/reference exists)"}
    cache = os.path.join(tempfile.gettempdir(), "pykmer_b200_reference_python_config1.json")
    if os.path.exists(cache):
        try:
            return json.load(open(cache))
        except Exception:
            pass
    import hashlib
    from pykmer_b200 import synth
    work = tempfile.mkdtemp(prefix="pykmer_ref_")
    try:
        src = os.path.join(work, "syn10M.fa.bgz")
        synth.write_fasta(src, synth.syn10m_records(), line_width=60, level=1)
        env = dict(os.environ, PYKMER_REFERENCE=ref)
        t0 = time.perf_counter()
        res = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_shim.py"), "indexer", src,
                              "syn10M", "11"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                             text=True, timeout=900)
        dt = time.perf_counter() - t0
        if res.returncode != 0:
            return {"unavailable": "the reference failed: " + res.stderr.strip().splitlines()[-1][:200]}
        meta = json.load(open(src + ".11.kin.json"))
        sha = hashlib.sha256(open(src + ".11.kin", "rb").read()).hexdigest()
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "indexer", "syn10M.fa.bgz.11.json")))
        bp = int(sum(synth.SYN10M_LENGTHS))
        out = {"value": bp / dt, "unit": "bp/s", "cores": 1, "kind": "reference", "seconds": dt,
               "sample": f"BASELINE config 1 in full: indexer.py syn10M.fa.bgz syn10M 11 ({bp} bp + 3 degenerate "
                         f"records, 4 MiB table), unmodified reference under CPython {sys.version.split()[0]}, "
                         f"wall time of the whole process",
               "num_kmers": meta["num_kmers"],
               "kin_matches_golden": bool(sha == gold["output_file_cheksum"])}
        with open(cache + ".tmp", "w") as fh:
            json.dump(out, fh)
        os.replace(cache + ".tmp", cache)
        return out
    except Exception as exc:                                   # never take the bench line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    finally:
        import shutil
        shutil.rmtree(work, ignore_errors=True)


def reference_indexer_line(args, K, stream, L, steps, warmup):
    table = np.zeros(4 ** K, dtype=np.uint8)
    vals, walls = [], []
    for it in range(warmup + steps):
        v, cores, what, wall = cpu_indexer_sample(stream, K, args.cpu_sample_mbp, table)
        if it >= warmup:
            vals.append(v); walls.append(wall)
    value = len(vals) / sum(1.0 / v for v in vals)          # total bp / total time
    name, unit = METRIC["indexer"]
    return {
        "impl": "reference", "metric": name.format(K=K), "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name("indexer", K, bp=L),
                   "note": "CPU arm: multithreaded C port of the reference algorithm (oracle/kmer_oracle.c), every "
                           "step the whole stream; the reference itself is single-threaded pure Python -- see "
                           "cpu_baseline.reference_python for it timed on config 1"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def reference_merger_line(args, K, N, max_count, steps, warmup):
    vals = []
    t0 = time.perf_counter()
    for it in range(warmup + steps):
        v, cores, what, _ = cpu_merger_sample(K, N, max_count)
        if it >= warmup:
            vals.append(v)
    wall = time.perf_counter() - t0
    value = len(vals) / sum(1.0 / v for v in vals)
    name, unit = METRIC["merger"]
    return {
        "impl": "reference", "metric": name, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * wall / (warmup + steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name("merger", K, N=N, max_count=max_count),
                   "note": "pair loop of the reference (C port, tools.py:473-482 per pair), extrapolated from a "
                           "bounded sample"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    if args.workload == "merger":
        print(json.dumps(reference_merger_line(args, args.kmer, args.samples, args.max_count, args.steps,
                                               args.warmup)), flush=True)
        return
    K = args.kmer
    stream, starts, lengths = load_stream(args.scale, 0, 1)
    line = reference_indexer_line(args, K, stream, int(sum(lengths)), args.steps, args.warmup)
    if args.workload == "all":
        sub = max(1, min(args.sub_steps, args.steps))
        line["merger"] = {"n50": reference_merger_line(args, 15, 50, 50, sub, 0),
                          "n255": reference_merger_line(args, 15, 255, 255, sub, 0)}
        if not args.no_reference_python:
            line["cpu_baseline"]["reference_python"] = reference_python()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm

def timed_steps(torch, dist, world, warmup, steps, body, sampler=None):
    """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events on the
    launching stream, max over ranks.  `sampler` keeps clock samples of the timed steps only."""
    for _ in range(warmup):
        body()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.timed = True
    ev0.record()
    for _ in range(steps):
        body()
    ev1.record()
    torch.cuda.synchronize()
    if sampler is not None and sampler.thread is not None:
        sampler.timed = False
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def wall_steps(torch, dist, world, warmup, steps, body):
    """The same bracket for steps that synchronise on the host themselves (pk_merge_host): wall
    clock between barriers, max over ranks."""
    for _ in range(warmup):
        body()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        body()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_indexer_seqshard(args, K, rank, local_rank, world, steps, warmup):
    """N > 1: sequence-sharded indexing.  Every rank scans 1/N of the stream
    with a scan-only handle over the full k-mer range; pass 2 stores the bucketed k-mer entries
    straight into the buffer of the rank that owns their table window (NVLink); each rank counts and
    commits its own windows; statistics are all-gathered."""
    import torch
    import torch.distributed as dist
    from pykmer_b200 import device as dev, dist as pdist, _native as nat

    T = 4 ** K
    stream, starts, lengths = load_stream(args.scale, rank, world)
    L = int(sum(lengths))
    a, b = pdist.slice_bounds(stream.size, rank, world)
    d_slice = torch.from_numpy(np.ascontiguousarray(stream[a:b])).cuda()
    halo = torch.from_numpy(np.ascontiguousarray(stream[max(0, a - 32):a])).cuda() if a > 0 else None
    scanner = dev.Indexer(K, device=local_rank, mode=nat.PK_MODE_SCAN)
    scanner.set_records(starts)

    def scan(src_host=None):
        scanner.reset()
        scanner.prime(halo, a)
        if src_host is None:
            scanner.feed_device(d_slice)
        else:
            scanner.feed_host(src_host)

    # window ownership balanced on the real k-mer distribution (one untimed scan)
    scan()
    all_cnt = pdist.gather_window_counts(scanner)
    owners = pdist.balanced_window_owners(all_cnt.sum(axis=(0, 1)), world)
    w0, w1 = owners[rank]
    wl = scanner.window_log2()
    lo, hi = w0 << wl, min(T, w1 << wl)
    counter = dev.Indexer(K, device=local_rank, range_lo=lo, range_hi=hi, mode=nat.PK_MODE_PARTITION)
    routed = args.exchange == "routed"
    fused = args.exchange in ("fused", "routed")
    if fused:
        pdist.connect_peer_pools(scanner, counter)
    if routed:
        pdist.setup_routed(scanner, counter, all_cnt.sum(axis=1), owners)
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    stage = torch.empty(b - a, dtype=torch.uint8, device="cuda") if fused else None
    last = {"exact_redo": 0}

    def step_device(src_host=None, table_out=None, exact=False):
        if fused:
            scanner.reset()
            scanner.prime(halo, a)
            counter.reset()
            src = d_slice
            if src_host is not None:
                stage.copy_(src_host, non_blocking=True)
                src = stage
            every = None
            if routed and not exact:
                status.zero_()
                every = pdist.exchange_routed(scanner, counter, src, status)
                buf = None
            else:
                buf = pdist.exchange_fused(scanner, counter, src, owners)
        else:
            scan(src_host)
            counter.reset()
            buf = pdist.exchange_entries(scanner, counter, owners)
        hist, st = counter.finalize(table_out=table_out)
        overflow = False
        if fused and routed and not exact:
            overflow, _, per_rank_kmers = pdist.read_routed_status(every)
            st["num_kmers"] = per_rank_kmers[rank]
        else:
            st["num_kmers"] = scanner.scan_result()
        hist, st = pdist.reduce_index_stats(hist, st)
        if overflow:
            # a region overflowed somewhere (the stream no longer looks like the planning scan): the
            # step is void, redo it with the exact two-pass protocol -- inside the timed region
            last["exact_redo"] += 1
            return step_device(src_host, table_out, exact=True)
        last["hist"], last["st"], last["buf"] = hist, st, buf

    sampler = ClockSampler(local_rank)
    l0 = scanner.launch_count() + counter.launch_count()
    sampler.start()
    ms = timed_steps(torch, dist, world, warmup, steps, step_device, sampler)
    clocks = sampler.stop()
    launches = (scanner.launch_count() + counter.launch_count() - l0) * steps // (steps + warmup)
    ms_step = ms / steps
    value = L / (ms_step * 1e-3)
    st = last["st"]

    scanner.set_profiling(True); counter.set_profiling(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if fused:
        scanner.reset(); scanner.prime(halo, a); counter.reset()
        ev[0].record()
        if routed:
            status.zero_()
            pdist.exchange_routed(scanner, counter, d_slice, status)
        else:
            pdist.exchange_fused(scanner, counter, d_slice, owners)
        ev[1].record()
    else:
        scan(); counter.reset()
        ev[0].record(); last["buf"] = pdist.exchange_entries(scanner, counter, owners); ev[1].record()
    counter.finalize()
    torch.cuda.synchronize()
    prof = dict(scanner.profile()); prof.update(counter.profile())
    scanner.set_profiling(False); counter.set_profiling(False)
    n_k_local = int(all_cnt[:, :, w0:w1].sum())
    table_bytes = hi - lo
    # every rank's share, for reading the max-over-ranks step time
    mine = torch.tensor([w1 - w0, n_k_local / 1e6, prof.get("scan_scatter", (0.0, 0))[0],
                         prof.get("window_count", (0.0, 0))[0], prof.get("window_commit", (0.0, 0))[0]],
                        dtype=torch.float64, device="cuda")
    every = torch.empty((world, 5), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(every.view(-1), mine)
    per_rank = [{"windows": int(a_), "entries_M": round(b_, 1), "scatter_ms": round(c_, 3), "count_ms": round(d_, 3),
                 "commit_ms": round(e_, 3)} for a_, b_, c_, d_, e_ in every.cpu().tolist()]
    peaks = measured_peaks()
    step_alg = stream.size + 64 * st["num_kmers"] + 2 * T
    # the byte-bound classes only: pass 2 is issue-bound and has no meaning as a fraction of HBM
    alg = {"window_count": 64 * n_k_local, "window_commit": 2 * table_bytes}
    dom = max((c for c in prof if c in alg), key=lambda c: prof[c][0])
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": alg[dom] / (prof[dom][0] * 1e-3) / 1e9,
                "peak": peaks["hbm"], "unit": "GB/s", "frac": alg[dom] / (prof[dom][0] * 1e-3) / 1e9 / peaks["hbm"],
                "traffic": None, "peak_source": peaks["hbm_src"], "rank": 0,
                "kernel_ms_by_class": {c: round(v[0], 4) for c, v in prof.items()},
                "exchange_ms_profiling_pass": round(ev[0].elapsed_time(ev[1]), 4),
                "exchange_bytes_sent": int(4 * (all_cnt[rank].sum() - all_cnt[rank, :, w0:w1].sum())),
                "step_algorithmic_bytes": step_alg,
                "step_frac": step_alg / (ms_step * 1e-3) / 1e9 / (peaks["hbm"] * world), "owners": owners,
                "note": "rank 0's kernels; k_scan_scatter is issue-bound (no byte roofline); step_frac is the whole "
                        "job's SURVEY 8d bytes against N x the measured HBM peak"}

    e2e = None
    if not args.no_e2e:
        h_slice = dev.pinned_empty(b - a)
        h_slice.numpy()[:] = stream[a:b]
        h_table = dev.pinned_empty(table_bytes)
        reps = max(2, min(steps, 3))
        ms_e = timed_steps(torch, dist, world, 2, reps, lambda: step_device(h_slice, h_table)) / reps
        e2e = {"value": L / (ms_e * 1e-3), "unit": "bp/s", "h2d_bytes_per_step": int(stream.size),
               "ms_per_step": ms_e, **e2e_transfer(torch, dist, world, counter, 257 * 8)}
        del h_table, h_slice
    flags = pdist.reduce_flags(scanner.record_flags())
    scanner.close(); counter.close()
    if rank != 0:
        return None
    name, unit = METRIC["indexer"]
    return {
        "metric": name.format(K=K), "value": value, "unit": unit, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name("indexer", K, bp=L),
                   "parallelism": f"sequence x{world} scan, k-mer entries to window owners "
                                  f"({'stores over NVLink fused into the scan, fixed regions, no host round trip' if routed else 'stores over NVLink fused into pass 2' if fused else 'NCCL all-to-all'}), "
                                  f"kmer-window x{world} count",
                   "exact_redo_steps": last["exact_redo"], "per_rank": per_rank,
                   "l2": L2_NOTE, "num_kmers": st["num_kmers"], "vals_sum": st["vals_sum"],
                   "vals_count": st["vals_count"], "vals_max": st["vals_max"],
                   "records_with_kmers": int(flags.sum())},
        "clocks": clocks, "roofline": roofline, "e2e": e2e, "gpu_launches": launches,
        "parity_check": golden_stats_check(K, args.scale, last["hist"], st),
    }


def e2e_transfer(torch, dist, world, ix, extra_bytes):
    """Bytes the last finalize(table_out=...) really moved device-to-host, summed over the ranks
    (pk_indexer_transfer_stats: sparse windows cross PCIe packed and are rebuilt on the host's cores
    inside the timed region), and how this rank's windows went."""
    x = ix.transfer_stats()
    total = x["d2h_bytes"] + extra_bytes
    if world > 1:
        t = torch.tensor([total], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total = int(t.item())
    return {"d2h_bytes_per_step": int(total),
            "d2h": {"packed_windows": x["packed_windows"], "raw_windows": x["raw_windows"],
                    "host_unpack_threads": x["unpack_threads"],
                    "note": "sparse table windows cross PCIe as bitmap + non-zero bytes (k_table_pack) and are rebuilt "
                            "in the caller's buffer by host threads inside the timed region; dense or backlogged "
                            "windows are copied as they are"}}


def golden_stats_check(K, scale, hist, st):
    """hist / vals_* / num_kmers of the step just timed against the oracle's digests of the same
    stream (tests/golden/at_scale.json, oracle/make_golden_at_scale.py)."""
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "at_scale.json")))
    except Exception:
        return None
    if scale != 1.0 or K not in (15, 17):
        return None
    parts = [gold["k15"]] if K == 15 else gold["k17"]
    want_hist = [sum(p["hist"][i] for p in parts) for i in range(255)]
    ok = list(hist) == want_hist and all(st[k] == sum(p[k] for p in parts)
                                         for k in ("num_kmers", "vals_sum", "vals_count"))
    return {"against": "oracle digests of the same stream (tests/golden/at_scale.json): hist[255], num_kmers, "
                       "vals_sum, vals_count", "equal": bool(ok)}


def bench_indexer(args, K, rank, local_rank, world, steps, warmup, cpu_sample_mbp=None):
    import torch
    import torch.distributed as dist
    from pykmer_b200 import device as dev

    T = 4 ** K
    stream, starts, lengths = load_stream(args.scale, rank, world)
    L = int(sum(lengths))
    from pykmer_b200 import dist as pdist
    lo, hi = pdist.shard_range(T, rank, world)               # k-mer-axis shard of this rank
    if args.emulate_shard:
        if ":" in args.emulate_shard:                        # "lo:hi" in GiB of the k-mer axis
            lo, hi = (int(float(x) * 2 ** 30) for x in args.emulate_shard.split(":"))
        else:
            er, en = (int(x) for x in args.emulate_shard.split("/"))
            lo, hi = pdist.shard_range(T, er, en)
    plan = None
    if world > 1 and K >= 19 and args.mode == 0:
        # very sparse tables count DIRECT: balance the shards on where the k-mers really fall.
        # One planning pass (each rank buckets 1/N of the stream per 2^26-entry window, the counts
        # are all-gathered), done once per genome and outside the timed steps like handle creation.
        from pykmer_b200 import _native as nat
        t0 = time.perf_counter()
        a, b = pdist.slice_bounds(stream.size, rank, world)
        scanner = dev.Indexer(K, device=local_rank, mode=nat.PK_MODE_SCAN)
        d_halo = torch.from_numpy(np.ascontiguousarray(stream[max(0, a - 32):a])).cuda() if a > 0 else None
        d_part = torch.from_numpy(np.ascontiguousarray(stream[a:b])).cuda()
        scanner.prime(d_halo, a)
        scanner.feed_device(d_part)
        per_window = pdist.gather_window_counts(scanner).sum(axis=(0, 1))      # synchronises
        del d_halo, d_part
        ranges = pdist.balanced_kmer_ranges(per_window, world, scanner.window_log2(), T)
        scanner.close()
        lo, hi = ranges[rank]
        plan = {"ranges_GiB": [round((h - l) / 2 ** 30, 2) for l, h in ranges],
                "plan_ms_untimed": round(1e3 * (time.perf_counter() - t0), 1)}

    d_stream = torch.from_numpy(stream).cuda()
    ix = dev.Indexer(K, device=local_rank, range_lo=lo, range_hi=hi, mode=args.mode)
    ix.set_records(starts)
    last = {}

    def step_device():
        ix.reset()
        ix.feed_device(d_stream)
        hist, st = ix.finalize()
        hist, st = pdist.reduce_index_stats(hist, st)         # one small NCCL all-gather
        last["hist"], last["st"] = hist, st

    sampler = ClockSampler(local_rank)
    launches0 = ix.launch_count()
    sampler.start()
    ms = timed_steps(torch, dist, world, warmup, steps, step_device, sampler)
    clocks = sampler.stop()
    launches_all = ix.launch_count() - launches0
    launches = launches_all * steps // (steps + warmup)
    ms_step = ms / steps
    value = L / (ms_step * 1e-3)
    st = last["st"]

    # per-kernel-class device time of one extra (untimed) step, CUDA events on the launching
    # stream around every launch (pk_indexer_set_profiling); the dominant class gets the roofline
    ix.set_profiling(True)
    ix.reset()
    ix.feed_device(d_stream)
    hist_l, st_l = ix.finalize()
    prof = ix.profile()
    ix.set_profiling(False)
    mode, windows = ix.mode()
    n_k_local = st_l["num_kmers"]
    table_bytes = hi - lo
    per_rank = None
    if world > 1:
        # every rank's share, for reading the max-over-ranks step time
        mine = torch.tensor([table_bytes / 2 ** 30, n_k_local / 1e6, sum(v[0] for v in prof.values())],
                            dtype=torch.float64, device="cuda")
        every = torch.empty((world, 3), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(every.view(-1), mine)
        per_rank = [{"table_GiB": round(a, 2), "num_kmers_M": round(b, 1), "kernel_ms": round(c, 3)}
                    for a, b, c in every.cpu().tolist()]
    big = table_bytes > 126e6
    # algorithmic bytes (SURVEY 8d / DESIGN.md): 1 B per base scanned, 64 B per counted k-mer
    # when the table exceeds L2 (one 32 B sector fetched + written back), 2 B per table entry
    alg_by_class = {
        "scan_count_direct": stream.size + (64 * n_k_local if big else 0),
        "window_count": 64 * n_k_local if big else 0,
        "window_commit": 2 * table_bytes,
        "table_stats": table_bytes,
    }
    step_alg = (stream.size + (64 * n_k_local if big else 0) + 2 * table_bytes) if world == 1 else \
        (stream.size * world + 64 * st["num_kmers"] + 2 * T)
    peaks = measured_peaks()
    peak = peaks["hbm"]
    dom = max((c for c in prof if c in alg_by_class), key=lambda c: prof[c][0])
    dom_ms, dom_launches = prof[dom]
    achieved = alg_by_class[dom] / (dom_ms * 1e-3) / 1e9
    kname = "k_window_count8" if (K >= 17 and dom == "window_count") else "k_" + dom
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kname) if (K in (15, 17) and world == 1 and not args.emulate_shard) else None,
                "peak_source": peaks["hbm_src"],
                "kernel_ms_total": dom_ms, "kernel_launches": dom_launches,
                "algorithmic_bytes": alg_by_class[dom],
                "algorithmic_bytes_per_launch": alg_by_class[dom] / max(dom_launches, 1),
                "counted_kmers_per_s": n_k_local / (dom_ms * 1e-3) if dom in ("window_count", "scan_count_direct") else None,
                "step_algorithmic_bytes": step_alg,
                "step_frac": step_alg / (ms_step * 1e-3) / 1e9 / (peak * world),
                "kernel_ms_by_class": {c: round(v[0], 4) for c, v in prof.items()},
                "mode": {1: "direct", 2: "partition"}.get(mode, str(mode)), "windows": windows}
    if dom in ("window_count", "scan_count_direct") and big:
        # the counting kernels are atomic-bound, not byte-bound: the second denominator is the random
        # atomic rate measured on this chip (tools/microbench*.cu, profiles/r01_microbench_b200.txt,
        # profiles/r01c_microbench2_b200.txt) for the operation the kernel issues
        op, peak_ops = (("red.add.u32, L2-resident 64 MiB window", 191.07e9) if (dom == "window_count" and K < 17) else
                        ("atom.add.u32 with return on packed bytes, L2-resident 64 MiB window", 128.0e9)
                        if dom == "window_count" else ("byte compare-and-swap, DRAM-resident table", 19.46e9))
        rate = n_k_local / (dom_ms * 1e-3)
        roofline["atomic"] = {"op": op, "achieved_per_s": rate, "peak_per_s": peak_ops, "frac": rate / peak_ops,
                              "peak_source": "builder-measured micro-benchmark on this pool's B200 (tools/microbench.cu, "
                                             "microbench2.cu), not a driver-measured peak"}
        if dom == "scan_count_direct" and K >= 19:
            roofline["note"] = ("step_frac may exceed 1: SURVEY 8d credits 2 B per table entry (zero-fill + statistics "
                                "pass); the DIRECT scheme writes the table once and keeps the histogram as transitions, "
                                "so it never reads the table back")
        if dom == "window_count":
            roofline["note"] = ("frac > 1 against HBM is not a physical HBM fraction: SURVEY 8d credits every counted "
                                "k-mer with 64 B of DRAM traffic (sector read + write-back), which counting in an "
                                "L2-resident window never performs; the honest figures are step_frac (whole step), "
                                "traffic (warm-cache ncu dram bytes per launch) and the atomic fraction")

    # end to end through the C ABI with HOST buffers: pinned stream in, table + stats out
    e2e = None
    if not args.no_e2e:
        h_stream = dev.pinned_empty(stream.size)
        h_stream.numpy()[:] = stream
        h_table = dev.pinned_empty(hi - lo)

        def step_e2e():
            ix.reset()
            ix.feed_host(h_stream)
            ix.finalize(table_out=h_table)

        reps = max(2, min(steps, 3))
        ms_e = timed_steps(torch, dist, world, 2, reps, step_e2e) / reps
        e2e = {"value": L / (ms_e * 1e-3), "unit": "bp/s", "h2d_bytes_per_step": int(stream.size),
               "ms_per_step": ms_e, **e2e_transfer(torch, dist, world, ix, 257 * 8)}
        if world == 1 and K == 15 and args.scale == 1.0 and not args.emulate_shard:
            # the table the last e2e step left in the HOST buffer against the oracle's digest of config 2
            try:
                gold = json.load(open(os.path.join(ROOT, "tests", "golden", "at_scale.json")))["k15"]["sha256"]
                e2e["host_table_sha256_equals_oracle"] = hashlib.sha256(h_table.numpy()).hexdigest() == gold
            except OSError:
                pass
        del h_table, h_stream

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, what, _ = cpu_indexer_sample(stream, K, args.cpu_sample_mbp if cpu_sample_mbp is None else cpu_sample_mbp)
        cpu = {"value": v, "unit": "bp/s", "cores": cores, "kind": "port", "sample": what,
               "reference_published": "503,287 bp/s (pypy, K=15, reference README.md:49)"}
    ix.close()
    del d_stream
    if rank != 0:
        return None
    name, unit = METRIC["indexer"]
    line = {
        "metric": name.format(K=K), "value": value, "unit": unit, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name("indexer", K, bp=L),
                   "parallelism": (f"kmer-range x{world}, replicated scan" + (", shards balanced on the k-mer distribution" if plan else ""))
                                  if world > 1 else
                                  (f"ONE shard ({args.emulate_shard}) of a k-mer-range job, development run"
                                   if args.emulate_shard else "single GPU"),
                   "l2": L2_NOTE, "num_kmers": st["num_kmers"], "vals_sum": st["vals_sum"],
                   "vals_count": st["vals_count"], "vals_max": st["vals_max"]},
        "clocks": clocks, "roofline": roofline, "e2e": e2e, "gpu_launches": launches,
        "parity_check": golden_stats_check(K, args.scale, last["hist"], st) if not args.emulate_shard else None,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if plan is not None:
        line["config"]["shard_plan"] = plan
    if per_rank is not None:
        line["config"]["per_rank"] = per_rank
    return line


def gram_issued_macs_per_step(N: int) -> int:
    """Multiply-accumulates the FP4 Gram kernel ISSUES per K=64 instruction step (gram_f4.cu): one
    128 x 64 tile for <= 64 samples, one 128 x 128 up to 128, three 128 x 128 ((0,0), (0,1), (1,1))
    up to 256 -- padding rows and, for <= 64, the unused half of the M=128 instruction included."""
    return 128 * 64 * 64 if N <= 64 else (128 * 128 * 64 if N <= 128 else 3 * 128 * 128 * 64)


_HOST_TABLES = {}


def bench_merger(args, K, N, max_count, rank, local_rank, world, steps, warmup):
    import torch
    import torch.distributed as dist
    from pykmer_b200 import device as dev

    T = 4 ** K
    from pykmer_b200 import dist as pdist
    lo, hi = pdist.shard_range(T, rank, world)
    n = hi - lo
    words = n // 32
    stride = (words + 3) & ~3
    tiled = dev.use_tiled_masks(N)        # default for <= 256 samples: tiled masks + the FP4 Gram kernel
    bits = dev.tiled_masks(words, N) if tiled else torch.zeros((N, stride), dtype=torch.int32, device="cuda")
    check = not args.no_parity_check and tiled
    rows = torch.zeros((N, stride), dtype=torch.int32, device="cuda") if check else None
    raw = torch.empty(n, dtype=torch.uint8, device="cuda")
    G = torch.zeros((N, N), dtype=torch.int64, device="cuda")
    # pack stage timed separately (per sample: generate, then threshold + pack)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    pack_ms = 0.0
    for s in range(N):
        dev.synth_table(s, lo, hi, out=raw)
        ev[0].record()
        if tiled:
            dev.threshold_pack_tiled(raw, 1, max_count, bits, s, N)
        else:
            dev.threshold_pack(raw, 1, max_count, out=bits[s])
        ev[1].record()
        torch.cuda.synchronize()
        pack_ms += ev[0].elapsed_time(ev[1])
        if check:
            dev.threshold_pack(raw, 1, max_count, out=rows[s])

    # the default path's matrix against the two other exact implementations on the same tables
    parity = None
    if check:
        G_f4 = dev.gram_tiled(bits, N, words).cpu().numpy()
        os.environ["PYKMER_B200_GRAM"] = "i8"
        try:
            G_i8 = dev.gram(rows, words=words).cpu().numpy()
            os.environ["PYKMER_B200_GRAM"] = "popc"
            n_sub = min(N, 50)
            G_pc = dev.gram(rows[:n_sub].contiguous(), words=words).cpu().numpy()
        finally:
            os.environ.pop("PYKMER_B200_GRAM", None)
        parity = {"against": f"k_gram_i8 (integer tensor-core accumulators) on row-major masks of the same tables, "
                             f"all {N * N} cells; k_gram_popc (AND + popcount) on the first {n_sub} samples",
                  "equal": bool(np.array_equal(G_f4, G_i8) and np.array_equal(G_f4[:n_sub, :n_sub], G_pc)),
                  "this_rank_only": world > 1}
    del rows

    def step():
        if tiled:
            dev.gram_tiled(bits, N, words, out=G, accumulate=False)
        else:
            dev.gram(bits, words=words, out=G, accumulate=False)
        pdist.reduce_gram(G)                                  # N x N int64 partials, NCCL all-reduce

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed_steps(torch, dist, world, warmup, steps, step, sampler)
    clocks = sampler.stop()
    ms_step = ms / steps
    bytes_bits = N * T / 8
    value = bytes_bits / (ms_step * 1e-3) / 1e9
    peaks = measured_peaks()
    algo = "f4" if tiled else os.environ.get("PYKMER_B200_GRAM", "i8" if N <= 256 else "popc")
    Gh = G.cpu().numpy()
    del bits

    # end to end through the C ABI with HOST tables (pk_merge_host): pinned copy in, pack, Gram
    e2e = None
    if not args.no_e2e:
        # every rank moves ITS slice of each of the N tables from pinned host memory.  The box cannot pin
        # N = 255 whole tables (274 GB): the first `distinct` samples are real and the rest alias them --
        # the same bytes cross PCIe and the same contraction runs, and the cells checked below belong
        # to real samples
        distinct = min(N, 16)
        key = (K, lo, hi)
        pool = _HOST_TABLES.setdefault(key, [])
        for s in range(len(pool), distinct):
            dev.synth_table(s, lo, hi, out=raw)
            h = dev.pinned_empty(n)
            h.copy_(raw)
            pool.append(h.numpy())
        host = [pool[s % distinct] for s in range(N)]
        torch.cuda.synchronize()
        reps, out = 2, {}
        Gp = torch.zeros((N, N), dtype=torch.int64, device="cuda")

        def step_e2e():
            m = dev.merge_host(host, 1, max_count, device=local_rank)
            Gp.copy_(torch.from_numpy(m[:, :, 2].astype(np.int64)))
            pdist.reduce_gram(Gp)
            out["G"] = Gp

        dt = wall_steps(torch, dist, world, 1, reps, step_e2e) / reps * 1e-3      # one untimed call first: allocator, exactness check
        Ge = out["G"].cpu().numpy()
        ok = int(Ge[0, 1]) == int(Gh[0, 1]) and int(Ge[0, 0]) == int(Gh[0, 0]) and \
            int(Ge[distinct - 1, 0]) == int(Gh[distinct - 1, 0])
        e2e = {"value": bytes_bits / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(N * T),
               "d2h_bytes_per_step": int(N * N * 3 * 8 * world), "ms_per_step": dt * 1e3,
               "table_bytes_per_s": N * T / dt, "distinct_host_tables": distinct,
               "matches_device_resident_result": bool(ok)}
        del host
    del raw

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, what, _ = cpu_merger_sample(K, N, max_count)
        cpu = {"value": v, "unit": "GB/s", "cores": cores, "kind": "port", "sample": what,
               "reference_published": "25.7 s per K=15 pair incl. gunzip (pypy, reference README.md:75-77)"}
    if rank != 0:
        return None
    name, unit = METRIC["merger"]
    hbm = {"floor_ms": bytes_bits / (peaks["hbm"] * 1e9 * world) * 1e3, "achieved_GBps": value,
           "peak_GBps": peaks["hbm"] * world, "frac": value / (peaks["hbm"] * world),
           "peak_source": peaks["hbm_src"]}
    if algo in ("i8", "f4"):
        # dense contraction on the tensor pipe.  `achieved` counts the operations the kernel ISSUES
        # (padding included, the symmetric form's 3 of 4 blocks); useful_* counts 2 N^2 4^K.
        # int8 runs at twice, FP4 at four times the bf16 rate.
        mult, nominal, kname = (2.0, 4500.0, "k_gram_i8") if algo == "i8" else \
            (4.0, 9000.0, "k_gram_f4" + (" (tiled masks)" if tiled else " (row-major masks)"))
        steps_k = T / 64.0
        issued = 2.0 * gram_issued_macs_per_step(N) * steps_k if algo == "f4" else 2.0 * N * N * T
        useful = 2.0 * N * N * T
        peak_t = mult * peaks["bf16"] * world                 # whole-job rate against the whole job's pipes
        ach = issued / (ms_step * 1e-3) / 1e12
        tensor_floor_ms = issued / (peak_t * 1e12) * 1e3
        roof = {"bound": "tensor" if tensor_floor_ms >= hbm["floor_ms"] else "hbm", "kernel": kname,
                "achieved": ach, "peak": peak_t, "unit": "TFLOP/s", "frac": ach / peak_t,
                "traffic": int(N * n / 8),
                "traffic_note": "ncu --set full of k_gram_f4 at K=13, N=255 (profiles/r01g_ncu_gram_f4_n255_k13.txt): "
                                "dram__bytes_read 2.140 GB = N * 4^13 / 8 exactly, every mask word is read once and "
                                "nothing is written; the figure here is that identity at this size, per launch",
                "peak_source": f"{mult:.0f} x {peaks['bf16_src']}" + (f" x {world} GPUs" if world > 1 else "")
                               + f"; nominal dense {nominal:.0f} TFLOP/s per GPU",
                "ops_counted": "issued by the kernel (gram_issued_macs_per_step: padding rows and the symmetric "
                               "form's 3 of 4 blocks as executed)",
                "frac_of_nominal": ach / (nominal * world),
                "useful_TFLOPs": useful / (ms_step * 1e-3) / 1e12, "useful_frac": useful / (ms_step * 1e-3) / 1e12 / peak_t,
                "tensor_floor_ms": tensor_floor_ms, "hbm": hbm}
        if roof["bound"] == "hbm":
            roof.update({"achieved": value, "peak": peaks["hbm"] * world, "unit": "GB/s", "frac": hbm["frac"],
                         "tensor": {"achieved_TFLOPs": ach, "peak_TFLOPs": peak_t, "frac": ach / peak_t}})
    else:
        popc = N * (N + 1) / 2 * T / 32
        roof = {"bound": "hbm", "kernel": "k_gram_popc", "achieved": value, "peak": peaks["hbm"],
                "unit": "GB/s", "frac": value / peaks["hbm"], "traffic": None, "peak_source": peaks["hbm_src"],
                "and_popc_per_s": popc / (ms_step * 1e-3), "and_popc_peak_measured": 4.378e12}
    line = {
        "metric": name, "value": value, "unit": unit, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": {"i8": "u8 x u8 -> s32 (tcgen05 kind::i8) -> int64",
                  "f4": "e2m1 x e2m1 -> f32 (tcgen05 kind::mxf4, exact integers < 2^24) -> int64"}.get(algo, "u32 popcount -> int64"),
        "data": "synthetic",
        "config": {"workload": workload_name("merger", K, N=N, max_count=max_count),
                   "parallelism": f"kmer-axis x{world}, one all-reduce of the N x N partial matrices" if world > 1 else "single GPU",
                   "l2": f"bitmask {bytes_bits / 1e9:.2f} GB >> 126 MB L2", "algo": algo,
                   "mask_layout": "tiled [1024 k-mers][sample][32 words]" if tiled else "row-major",
                   "pack_ms_total": pack_ms, "pack_GBps": N * n * 1.125 / (pack_ms * 1e-3) / 1e9,
                   "trace_G": int(np.trace(Gh)), "G01": int(Gh[0, 1])},
        "clocks": clocks, "roofline": roof, "e2e": e2e, "gpu_launches": steps,
        "parity_check": parity,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    return line


def free_device_memory():
    import gc
    import torch
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def indexer(K, steps, warmup, cpu_sample_mbp=None):
        shard = args.shard
        if shard == "auto":
            # very sparse tables (K >= 19: 16384 windows) are bound by per-window launches, where
            # the replicated scan with k-mer ranges measured faster (profiles/)
            shard = "sequence" if (4 ** K >> 24) <= 4096 else "kmer"
        if world > 1 and shard == "sequence":
            return bench_indexer_seqshard(args, K, rank, local_rank, world, steps, warmup)
        return bench_indexer(args, K, rank, local_rank, world, steps, warmup, cpu_sample_mbp)

    try:
        if args.workload == "merger":
            line = bench_merger(args, args.kmer, args.samples, args.max_count, rank, local_rank, world,
                                args.steps, args.warmup)
        else:
            line = indexer(args.kmer, args.steps, args.warmup)
        if args.workload == "all":
            sub, subw = max(1, min(args.sub_steps, args.steps)), max(1, min(3, args.warmup))
            extra = {}
            free_device_memory()
            extra["indexer_k17"] = indexer(17, sub, subw, cpu_sample_mbp=128.0)
            free_device_memory()
            if world >= 4:
                extra["indexer_k19"] = indexer(19, sub, subw)
                free_device_memory()
            merger = {}
            for tag, N, mc in (("n50", 50, 50), ("n255", 255, 255)):
                merger[tag] = bench_merger(args, 15, N, mc, rank, local_rank, world, sub, subw)
                free_device_memory()
            extra["merger"] = merger
            _HOST_TABLES.clear()
            if rank == 0:
                line.update(extra)
                if world == 1 and not args.no_cpu_baseline and not args.no_reference_python \
                        and "cpu_baseline" in line:
                    line["cpu_baseline"]["reference_python"] = reference_python()
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
