#!/usr/bin/env python3
"""Manufacture tests/golden/ by running the reference itself (build container only).

TEST INFRASTRUCTURE ONLY.  Runs /root/reference's indexer.py and merger.py
through oracle/ref_shim.py on small, committed inputs and stores what they
wrote: the .kin bytes (or their sha256 for the 10 Mbp case), the deterministic
keys of .kin.json, and the merger's matrix.  The committed fixtures are what
pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

    python oracle/make_golden.py            # regenerate everything (~1-2 min)
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykmer_b200 import synth  # noqa: E402  (input generators only)

GOLD = os.path.join(ROOT, "tests", "golden")
SHIM = os.path.join(ROOT, "oracle", "ref_shim.py")

DETERMINISTIC_KEYS = [
    "file_ver", "kmer_size", "data_size", "max_size", "project_name", "kmer_len",
    "flush_every", "frag_size", "input_file_name", "input_file_size", "input_file_cheksum",
    "output_file_size", "output_file_cheksum", "num_kmers", "chromosomes", "hist", "hist_sum",
    "hist_count", "hist_min", "hist_max", "vals_sum", "vals_count", "vals_min", "vals_max",
]


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run_ref(*args: str) -> None:
    subprocess.run([sys.executable, SHIM, *args], check=True, stdout=subprocess.DEVNULL)


# ------------------------------------------------------------------ hand-made inputs

def tiny_mixed() -> bytes:
    """Every text rule of indexer.py:45-99 and Appendix A of SURVEY.md in one file."""
    return (
        b"ACGTACGTACGTACGT text before the first header is dropped\n"
        b"\n"
        b">rec1 first record with a description\r\n"
        b"ACGTTGCAAGCTAGCTAGGATCCGATCGATTAGC\r\n"
        b"  acgtnnacgtacgtaacc  \n"
        b"\n"
        b"GGGGCCCCAAAATTTTRYKMACGTACGTAGCTAGCATCGA\n"
        b">rec2\n"
        b"ACG\n"
        b">rec3 empty record follows\n"
        b">rec4 interior blank resets the window\n"
        b"ACGTACGTAC GTACGTACGTAGCATGCATGCAT\n"
        b"\tTTTTTTTTTTTTTTTTTTTTTTTTGGGGG\t\n"
        b">rec5 only N\n"
        b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNN\n"
        b">rec1 first record with a description\n"
        b"ACGTTGCAAGCTAGCTAGGATCCGATCGATTAGC*-.acgt\n"
        b" > not a header after strip? yes it is: stripped line starts with >\n"
        b"TTGACCAGTAGGACCATTAGGACATTTAGGA\n"
        b">last record has no trailing newline\n"
        b"CATGCATGCATGCCCGGGTTTAAACATGCATGAAGGCCTT"
    )


def saturating() -> list:
    rng = np.random.default_rng(7)
    a = np.full(3000, ord("A"), dtype=np.uint8)
    at = np.frombuffer(b"AT" * 1200, dtype=np.uint8)
    aat = np.frombuffer(b"aat" * 700, dtype=np.uint8)
    rnd = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5000)]
    return [("polyA run", a), ("AT microsatellite", at), ("aat lower-case", aat),
            ("random tail", rnd), ("polyT is the same canonical k-mer as polyA",
                                   np.full(2000, ord("T"), dtype=np.uint8))]


def all_kmers(K: int) -> list:
    """test.py:8-27 construction: every K-string as its own record."""
    recs = []
    for v in range(4 ** K):
        s = bytes(b"ACGT"[(v >> (2 * (K - 1 - p))) & 3] for p in range(K))
        recs.append((f"examples/example--{K:02d}-{v + 1:010d}", np.frombuffer(s, dtype=np.uint8)))
    return recs


def write_inputs(d: str) -> dict:
    os.makedirs(d, exist_ok=True)
    files = {}
    with open(os.path.join(d, "tiny_mixed.fa"), "wb") as fh:
        fh.write(tiny_mixed())
    files["tiny_mixed.fa"] = [3, 5, 7, 9]
    synth.write_fasta(os.path.join(d, "saturating.fa.gz"), saturating(), line_width=70)
    files["saturating.fa.gz"] = [3, 5, 7, 11]
    synth.write_fasta(os.path.join(d, "allkmers_05.fasta.gz"), all_kmers(5), line_width=80)
    files["allkmers_05.fasta.gz"] = [5, 3]
    synth.write_fasta(os.path.join(d, "allkmers_07.fasta.gz"), all_kmers(7), line_width=80)
    files["allkmers_07.fasta.gz"] = [7]
    recs = synth.synth_genome(1234, (120_000, 60_000, 20_000), name_fmt="rand{:02d}",
                              library_elements=20)
    synth.write_fasta(os.path.join(d, "rand200k.fa.bgz"), recs, line_width=61, newline=b"\r\n")
    files["rand200k.fa.bgz"] = [9, 11, 13]
    return files


# ------------------------------------------------------------------ golden runs

def golden_indexer(inp_dir: str, files: dict, work: str) -> None:
    out_dir = os.path.join(GOLD, "indexer")
    os.makedirs(out_dir, exist_ok=True)
    for fname, ks in files.items():
        src = os.path.join(work, fname)
        shutil.copy(os.path.join(inp_dir, fname), src)
        for K in ks:
            run_ref("indexer", src, "sample", str(K))
            kin = f"{src}.{K:02d}.kin"
            with open(kin + ".json") as fh:
                meta = json.load(fh)
            keep = {k: meta[k] for k in DETERMINISTIC_KEYS}
            keep["project_name"] = os.path.basename(keep["project_name"])
            keep["all_keys"] = sorted(meta.keys())
            table = np.fromfile(kin, dtype=np.uint8)
            assert hashlib.sha256(table.tobytes()).hexdigest() == meta["output_file_cheksum"]
            stem = os.path.join(out_dir, f"{fname}.{K:02d}")
            with open(stem + ".json", "w") as fh:
                json.dump(keep, fh, indent=1, sort_keys=True)
            if K <= 11:
                np.savez_compressed(stem + ".kin.npz", table=table)
            print(f"  indexer {fname} K={K}: num_kmers={meta['num_kmers']} "
                  f"vals_max={meta['vals_max']}")


def golden_config1(work: str) -> None:
    """BASELINE config 1: 10 Mbp synthetic multi-FASTA (bgzip), K=11.  Only digests
    and statistics are committed; the FASTA is regenerated from its seed."""
    out_dir = os.path.join(GOLD, "indexer")
    src = os.path.join(work, "syn10M.fa.bgz")
    synth.write_fasta(src, synth.syn10m_records(), line_width=60, level=1)
    run_ref("indexer", src, "syn10M", "11")
    with open(src + ".11.kin.json") as fh:
        meta = json.load(fh)
    keep = {k: meta[k] for k in DETERMINISTIC_KEYS}
    keep["project_name"] = os.path.basename(keep["project_name"])
    keep["fasta_sha256"] = sha256(src)
    with open(os.path.join(out_dir, "syn10M.fa.bgz.11.json"), "w") as fh:
        json.dump(keep, fh, indent=1, sort_keys=True)
    print(f"  indexer syn10M K=11: num_kmers={meta['num_kmers']} vals_max={meta['vals_max']}")


def golden_merger(work: str) -> None:
    """Reference merger on (a) .kin files the reference indexer just wrote and
    (b) synthetic tables written in the reference's format, under several
    threshold pairs; one input is gzip-compressed (.kin.bgz, tools.py:296-302)."""
    out_dir = os.path.join(GOLD, "merger")
    os.makedirs(out_dir, exist_ok=True)
    K = 7
    mdir = os.path.join(work, "merge")
    os.makedirs(mdir, exist_ok=True)
    # (b) five synthetic samples + (a) three indexed ones at the same K
    rng = np.random.default_rng(99)
    kins = []
    template = json.load(open(os.path.join(work, "tiny_mixed.fa.07.kin.json")))
    for s in range(5):
        t = synth.synth_table(s, K)
        if s == 4:  # dense, wide-valued sample so every threshold bites
            t = rng.integers(0, 256, size=4 ** K, dtype=np.uint8)
        base = os.path.join(mdir, f"synth{s}.fa")
        open(base, "w").close()
        kin = f"{base}.{K:02d}.kin"
        t.tofile(kin)
        meta = dict(template)
        meta.update(input_file_name=os.path.basename(base), input_file_path=base,
                    project_name=base, kmer_len=K)
        with open(kin + ".json", "w") as fh:
            json.dump(meta, fh, indent=1, sort_keys=True)
        if s == 2:  # compressed input: merger must gunzip it
            with open(kin, "rb") as fi, gzip.open(kin + ".bgz", "wb") as fo:
                fo.write(fi.read())
            os.remove(kin)
            kin = kin + ".bgz"
        kins.append(kin)
    for f in ("tiny_mixed.fa", "saturating.fa.gz", "allkmers_07.fasta.gz"):
        kins.append(os.path.join(work, f"{f}.{K:02d}.kin"))
    kins_sorted = sorted(kins)
    tables = []
    for k in kins_sorted:
        if k.endswith(".bgz"):
            tables.append(np.frombuffer(gzip.open(k, "rb").read(), dtype=np.uint8))
        else:
            tables.append(np.fromfile(k, dtype=np.uint8))
    np.savez_compressed(os.path.join(out_dir, "samples_K07.npz"), tables=np.stack(tables),
                        names=np.array([os.path.basename(k) for k in kins_sorted]))
    for lo, hi in ((1, 255), (2, 10), (1, 50), (255, 255), (1, 1), (3, 3)):
        proj = os.path.join(mdir, f"proj_{lo}_{hi}")
        run_ref("merger", proj, *kins, f"--min-count={lo}", f"--max-count={hi}", "--threads=2",
                "--block-size=5000")
        kma = f"{proj}.{lo:03d}-{hi:03d}.kma"
        m = np.load(kma)["matrix"]
        desc = json.load(open(kma + ".json"))
        N = m.shape[0]
        m[np.arange(N), np.arange(N), :] = 0  # the reference leaves the diagonal uninitialised
        np.savez_compressed(os.path.join(out_dir, f"matrix_K07_{lo:03d}-{hi:03d}.npz"), matrix=m)
        slim = {"min_count": desc["min_count"], "max_count": desc["max_count"],
                "top_keys": sorted(desc.keys()),
                "data_keys": sorted(desc["data"][0].keys()),
                "header_keys": sorted(desc["data"][0]["header"].keys()),
                "order": [os.path.basename(d["index_file"]) for d in desc["data"]],
                "pos": [d["pos"] for d in desc["data"]]}
        with open(os.path.join(out_dir, f"matrix_K07_{lo:03d}-{hi:03d}.json"), "w") as fh:
            json.dump(slim, fh, indent=1, sort_keys=True)
        print(f"  merger K=7 [{lo},{hi}] N={N} shared[0,1]={int(m[0, 1, 2])}")


def main() -> None:
    inp_dir = os.path.join(GOLD, "inputs")
    files = write_inputs(inp_dir)
    work = tempfile.mkdtemp(prefix="pykmer_golden_")
    try:
        golden_indexer(inp_dir, files, work)
        golden_merger(work)
        if "--skip-config1" not in sys.argv:
            golden_config1(work)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
