#!/usr/bin/env python3
"""Wall time of the drop-in CLI on BASELINE config 2: writes the synthetic 782.5 Mbp genome as a
bgzip multi-FASTA (cached in the temp dir), runs `indexer.py <file> <sample> 15`, prints one JSON
line.  (The reference needs 27 min 50 s for this under pypy, reference README.md:49.)
    python tools/cli_e2e.py [scale] [K]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pykmer_b200 import indexer, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
K = int(sys.argv[2]) if len(sys.argv) > 2 else 15
path = os.path.join(tempfile.gettempdir(), f"syn782M_{scale:.4f}.fa.bgz")
if not os.path.exists(path):
    t0 = time.perf_counter()
    synth.write_fasta(path + ".tmp.bgz", synth.syn782m_records(scale=scale), line_width=60, level=1)
    os.replace(path + ".tmp.bgz", path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.0f} MB) in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
for rep in range(2):                      # second run: warm page cache and CUDA context
    t0 = time.perf_counter()
    header = indexer.create_fasta_index(path, "syn782M", path, K, overwrite=True, buffer_size=2 ** 16)
    wall = time.perf_counter() - t0
bp = sum(l for _, l in header.chromosomes)
print(json.dumps({"cli": f"indexer.py syn782M.fa.bgz syn782M {K}", "bases": bp, "wall_s": wall,
                  "bp_per_s": bp / wall, "phases_s": header.wall_seconds, "num_kmers": header.num_kmers,
                  "vals_count": header.vals_count, "cores": os.cpu_count(),
                  "reference_pypy_wall_s": 1670 if K == 15 else None}))
