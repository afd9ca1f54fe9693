#!/usr/bin/env python3
"""Drop-in for the reference's test.py (test.py:8-33): write examples/example--KK.fasta.gz holding
every one of the 4^K strings over ACGT as its own record -- the only known-answer input the
reference ships.  Indexed at K it must give num_kmers = 4^K and table[i] = 2 iff i <= rc(i)
(tests/test_oracle_golden.py, tests/test_gpu_parity.py check exactly that).

    test.py [K ...]        default 3 5 7 9 11 (the reference's list goes on to 21: 4^21 records)

Same file names (double dash included, test.py:21,32), record names
`examples/example--KK-<10-digit number from 1>`, record order (A < C < G < T, first base most
significant) and one sequence line per record; existing files are kept (test.py:23).
"""
import gzip
import os
import sys

import numpy as np

ALFA = np.frombuffer(b"ACGT", dtype=np.uint8)


def create_test(seq_name: str, kmer_len: int, batch: int = 1 << 16) -> str:
    fasta_file = f"{seq_name}-{kmer_len:02d}.fasta.gz"
    if os.path.exists(fasta_file):
        return fasta_file
    os.makedirs(os.path.dirname(fasta_file) or ".", exist_ok=True)
    shifts = 2 * (kmer_len - 1 - np.arange(kmer_len, dtype=np.uint64))
    with gzip.open(fasta_file, "wb") as fhd:
        for lo in range(0, 4 ** kmer_len, batch):
            vals = np.arange(lo, min(lo + batch, 4 ** kmer_len), dtype=np.uint64)
            seqs = ALFA[((vals[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.intp)]
            fhd.write(b"".join(b">%s-%02d-%010d\n%s\n" % (seq_name.encode(), kmer_len, int(v) + 1, s.tobytes())
                               for v, s in zip(vals, seqs)))
    return fasta_file


def main() -> None:
    ks = [int(a) for a in sys.argv[1:]] or [3, 5, 7, 9, 11]
    for kmer_len in ks:
        print(kmer_len)
        create_test("examples/example-", kmer_len)


if __name__ == "__main__":
    main()
