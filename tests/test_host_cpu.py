"""CPU-only checks: host logic, text rules, file formats, and that the C ABI loads
and exports every symbol include/pykmer_b200.h declares (no compute calls)."""
import glob
import gzip
import json
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle
from pykmer_b200 import fasta, synth
from pykmer_b200.tools import Header, Timer, frag_size_rule, gen_checksum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_kmer_bits_algebra_on_host(tmp_path):
    """The exact bit algebra the CUDA scan uses (kmer_bits.h), checked against a
    literal restatement of indexer.py:141-150 for every odd K <= 31."""
    exe = str(tmp_path / "test_kmer_bits")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-I", os.path.join(ROOT, "pykmer_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "test_kmer_bits.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.startswith("ok:")


def test_c_abi_exports_every_declared_symbol():
    from pykmer_b200 import _native as nat
    text = open(os.path.join(ROOT, "include", "pykmer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", text))
    assert declared == set(nat.SYMBOLS)
    for name in declared:
        assert getattr(nat.lib, name) is not None
    assert nat.lib.pk_abi_version() == 1


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "inputs", "*"))),
                         ids=lambda p: os.path.basename(p))
@pytest.mark.parametrize("chunk", [64 << 20, 4096, 61, 5])
@pytest.mark.parametrize("native", [True, False], ids=["native", "python"])
def test_fasta_reader_matches_reference_text_rules(path, chunk, native):
    recs = list(oracle.parse_records(path))
    want, starts, lengths, names = oracle.records_to_stream(recs)
    fs = fasta.FastaStream(path, chunk_bytes=chunk, native=native)
    got = fs.read_all()
    assert fs.names == names and fs.lengths == lengths and fs.starts == starts.tolist()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("text", [
    b"", b"\n\n", b"ACGT\n", b">only header", b">a\n>b\n>c\nAC\n", b">a\rACGT\rAC\r>b\r\rGG",
    b">a\r\nAC GT\r\n\tACGT \r\n", b"  >lead\n  ACGT\n", b">x\x0c\nAC\x0bGT\x1c\n",
    b">caf\xc3\xa9 name  \nACGT\n", b">a\nACGT>notheader\n>b\nTT\n",
])
def test_fasta_reader_edge_cases(tmp_path, text):
    p = str(tmp_path / "e.fa")
    open(p, "wb").write(text)
    recs = list(oracle.parse_records(p))
    want, starts, lengths, names = oracle.records_to_stream(recs)
    for chunk, native in ((1 << 20, True), (3, True), (1 << 20, False), (3, False)):
        fs = fasta.FastaStream(p, chunk_bytes=chunk, native=native)
        got = fs.read_all()
        assert fs.names == names and fs.lengths == lengths and fs.starts == starts.tolist()
        assert np.array_equal(got, want)


@pytest.mark.parametrize("native", [True, False], ids=["native", "python"])
def test_fasta_reader_rejects_non_ascii_sequence(tmp_path, native):
    p = str(tmp_path / "bad.fa")
    open(p, "wb").write(b">a\nAC\xc3\xa9GT\n")
    with pytest.raises(ValueError):
        fasta.FastaStream(p, native=native).read_all()


def test_native_ingest_equals_python_ingest_on_messy_bgzf(tmp_path):
    """csrc/ingest.cpp (parallel BGZF inflate + newline stripping) against this module's own
    Python path: CRLF / lone CR / blank lines / tabs / hidden headers / '>' inside sequence, several
    thread counts and chunk sizes, pieces written into a caller's buffer ring."""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    lines = [b"junk before the first header", b""]
    for r in range(40):
        lines.append(b">rec%d some description " % r)
        for _ in range(int(rng.integers(0, 400))):
            body = acgt[rng.integers(0, len(acgt), size=int(rng.integers(0, 90)))].tobytes()
            roll = rng.random()
            if roll < 0.02:
                body = b"  " + body + b"\t"
            elif roll < 0.03:
                body = body[:5] + b">" + body[5:]
            elif roll < 0.035:
                body = b" >hidden%d" % r
            lines.append(body)
    eols = [b"\n", b"\r\n", b"\r"]
    text = b"".join(l + eols[int(rng.integers(0, 3))] for l in lines)
    path = str(tmp_path / "messy.fa.bgz")
    open(path, "wb").write(synth.bgzf_compress(text, level=1))
    ref = fasta.FastaStream(path, native=False)
    want = ref.read_all()
    assert len(ref.names) > 40                             # hidden headers were found
    for chunk, threads in ((64 << 20, 0), (1 << 16, 3), (7001, 1)):
        fs = fasta.FastaStream(path, chunk_bytes=chunk, threads=threads, native=True)
        ring = [np.empty(max(chunk, 1 << 16) * 2 + 64, dtype=np.uint8) for _ in range(2)]
        got = np.concatenate([p.copy() for p in fs.pieces(buffers=ring)])
        assert np.array_equal(got, want)
        assert fs.names == ref.names and fs.starts == ref.starts and fs.lengths == ref.lengths
    plain = str(tmp_path / "messy.fa")
    open(plain, "wb").write(text)
    assert np.array_equal(fasta.FastaStream(plain, chunk_bytes=5000, native=True).read_all(), want)
    gz = str(tmp_path / "messy.fa.gz")                     # ordinary gzip: not BGZF
    open(gz, "wb").write(gzip.compress(text))
    assert np.array_equal(fasta.FastaStream(gz, native=True).read_all(), want)


def test_native_bgzf_inflate_rejects_damage(tmp_path):
    raw = b">a\n" + b"ACGT" * 100_000 + b"\n"
    blob = bytearray(synth.bgzf_compress(raw, level=1))
    good = str(tmp_path / "ok.fa.bgz")
    open(good, "wb").write(bytes(blob))
    assert fasta.FastaStream(good, native=True).read_all().size == 400_001
    blob[len(blob) // 2] ^= 0x55                           # flip bits inside a deflate stream
    bad = str(tmp_path / "bad.fa.bgz")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises((OSError, ValueError)):
        fasta.FastaStream(bad, native=True).read_all()
    cut = str(tmp_path / "cut.fa.bgz")
    open(cut, "wb").write(synth.bgzf_compress(raw, level=1)[:-40])
    with pytest.raises(OSError):
        fasta.FastaStream(cut, native=True).read_all()


def test_bgzf_writer_is_readable_as_gzip(tmp_path):
    raw = os.urandom(200_000) + b"ACGT" * 50_000
    blob = synth.bgzf_compress(raw, level=1)
    assert gzip.decompress(blob) == raw
    assert blob.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def test_header_names_sizes_and_frag_rule(tmp_path):
    f = str(tmp_path / "g.fa.gz")
    h = Header("proj", input_file=f, kmer_len=15)
    assert h.index_file_root == f + ".15.kin"
    assert h.index_tmp_file == f + ".15.kin.tmp" and h.metadata_file == f + ".15.kin.json"
    assert h.kmer_size == h.data_size == h.max_size == 4 ** 15
    assert h.file_ver == "KMER001" and h.max_val == 255 and h.frag_size == 357_914_000
    open(f + ".15.kin.bgz", "wb").close()
    assert h.index_file == f + ".15.kin.bgz"        # tools.py:185-190 prefers the .bgz
    for K in (3, 5, 7, 9, 11, 13, 15, 17, 19):
        assert frag_size_rule(4 ** K, 500_000_000, 1_000_000_000) == oracle.frag_size_rule(K)
    for bad in (0, 4, -1):
        with pytest.raises(AssertionError):        # tools.py:165-167
            Header("p", input_file=f, kmer_len=bad)


def test_header_metadata_roundtrip_and_errors(tmp_path):
    gold = json.load(open(os.path.join(GOLD, "indexer", "tiny_mixed.fa.05.json")))
    base = str(tmp_path / "tiny_mixed.fa")
    open(base, "w").close()
    meta = {k: gold.get(k) for k in Header.HEADER_FIXED + Header.HEADER_DATA}
    meta.update(input_file_path=base, kmer_len=5)
    json.dump(meta, open(base + ".05.kin.json", "w"))
    h = Header("whatever", index_file=base + ".05.kin")
    assert h.kmer_len == 5 and h.input_file_path == base and h.num_kmers == gold["num_kmers"]
    lean = h.to_dict(lean=True)
    assert "chromosomes" not in lean and len(lean) == 32 and len(h.to_dict()) == 33
    h2 = Header("whatever", index_file=base + ".05.kin.bgz")     # .bgz suffix is stripped
    assert h2.kmer_len == 5
    del meta["hist"]
    json.dump(meta, open(base + ".05.kin.json", "w"))
    with pytest.raises(KeyError):                                 # tools.py:394
        Header("whatever", index_file=base + ".05.kin")
    meta["hist"] = gold["hist"]
    meta["kmer_size"] = 7
    json.dump(meta, open(base + ".05.kin.json", "w"))
    with pytest.raises(AssertionError):                           # tools.py:401
        Header("whatever", index_file=base + ".05.kin")


def test_init_tmp_file_is_sparse_and_overwrite_rule(tmp_path):
    f = str(tmp_path / "x.fa")
    h = Header("p", input_file=f, kmer_len=5)
    h.init_index_tmp_file(overwrite=True)
    assert os.path.getsize(h.index_tmp_file) == 4 ** 5
    open(h.index_file_root, "wb").close()
    with pytest.raises(ValueError):                               # tools.py:319,325
        h.init_index_tmp_file(overwrite=False)
    h.init_index_tmp_file(overwrite=True)
    assert not os.path.exists(h.index_file_root)


def test_header_memmap_views_and_byte_iterator(tmp_path):
    """tools.py:240-243,344-363,527-533: init_index_file, the memmap accessors and __iter__
    (which inflates a .kin.bgz on the way, tools.py:296-302)."""
    from pykmer_b200 import bgzf
    K = 5
    f = str(tmp_path / "y.fa")
    h = Header("p", input_file=f, kmer_len=K)
    h.init_index_file(overwrite=True)
    assert os.path.getsize(h.index_file_root) == 4 ** K
    table = synth.synth_table(1, K)
    for arr in h.get_array_from_index_file():
        assert arr.shape == (4 ** K,) and arr.dtype == np.uint8 and not arr.any()
        arr[:] = table
        arr.flush()
    assert np.array_equal(np.fromfile(h.index_file_root, dtype=np.uint8), table)
    assert list(h) == table.tolist()
    h.init_file(h.index_tmp_file)
    for arr in h.get_array_from_index_tmp_file(fhd_mode="rb", mm_mode="r"):
        assert arr.size == 4 ** K and not arr.any()
    bgzf.compress_file(h.index_file_root, level=6)
    assert h.index_file.endswith(".kin.bgz") and list(h) == table.tolist()


def test_timer_and_checksum(tmp_path):
    t = Timer()
    t.update(1000)
    assert t.val_last == 1000 and t.speed_ela >= 0
    p = tmp_path / "c.bin"
    p.write_bytes(b"abc")
    assert gen_checksum(str(p)) == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"


def test_synth_tables_are_slice_consistent():
    full = synth.synth_table(3, 7)
    assert np.array_equal(full[1000:5000], synth.synth_table_slice(3, 1000, 5000))
    nz = np.count_nonzero(synth.synth_table(0, 9)) / 4 ** 9
    assert 0.12 < nz < 0.20
    assert synth.synth_table(1, 9).max() == 255


def test_merger_argument_validation(tmp_path):
    from pykmer_b200 import merger
    with pytest.raises(AssertionError):
        merger.merge(str(tmp_path / "p"), [tmp_path / "a.kin"], min_count=0)
    with pytest.raises(AssertionError):
        merger.merge(str(tmp_path / "p"), [tmp_path / "a.kin"], max_count=256)
    with pytest.raises(AssertionError):                           # inputs must exist
        merger.merge(str(tmp_path / "p"), [tmp_path / "a.kin", tmp_path / "b.kin"])
    (tmp_path / "a.txt").write_bytes(b"")
    with pytest.raises(AssertionError):                           # extension rule merger.py:109
        merger.merge(str(tmp_path / "p"), [tmp_path / "a.txt"])
    with pytest.raises(SystemExit):                               # merger.py:224-226
        merger.main([str(tmp_path / "p"), str(tmp_path / "a.kin")])
    args = merger.build_parser().parse_args(["proj", "a.kin", "b.kin", "--max-count=50"])
    assert args.max_count == 50 and args.min_count == 1 and args.threads == 4


def test_bgzf_module_roundtrip_and_kin_bgz_reader(tmp_path):
    """.kin.bgz written by the parallel BGZF writer: gzip reads it (what the reference's merger does,
    tools.py:296-302), the parallel reader reads it, Header.read_table takes either flavour."""
    from pykmer_b200 import bgzf
    K = 7
    table = synth.synth_table(2, K)
    base = str(tmp_path / "s.fa")
    kin = base + ".07.kin"
    table.tofile(kin)
    out = bgzf.compress_file(kin, level=9, threads=3, batch=2)
    assert out == kin + ".bgz" and fasta.is_bgzf(out)
    assert gzip.open(out, "rb").read() == table.tobytes() == bgzf.read_all(out)
    h = Header("p", input_file=base, kmer_len=K)
    assert h.index_file == out                                   # .bgz preferred (tools.py:185-190)
    assert np.array_equal(h.read_table(), table)
    os.remove(out)
    with gzip.open(out, "wb") as fz:                             # plain gzip under the .bgz name
        fz.write(table.tobytes())
    assert not fasta.is_bgzf(out) and np.array_equal(h.read_table(), table)
    assert bgzf.decompress_file(bgzf.compress_file(kin, dst=str(tmp_path / "x.bgz")), str(tmp_path / "x.kin"))
    assert open(tmp_path / "x.kin", "rb").read() == table.tobytes()


def test_bgz_table_roundtrip_through_native_reader(tmp_path):
    """.kin -> bgzf.compress_file -> Header.read_table (pk_bgzf_inflate straight into the table array)
    and the block-parallel Python reader give the bytes back; a table of the wrong size is refused."""
    from pykmer_b200 import bgzf
    K = 9
    rng = np.random.default_rng(3)
    table = (rng.random(4 ** K) < 0.2).astype(np.uint8) * rng.integers(1, 256, size=4 ** K).astype(np.uint8)
    fa = str(tmp_path / "s.fa")
    open(fa, "w").close()
    kin = fa + f".{K:02d}.kin"
    table.tofile(kin)
    bgz = bgzf.compress_file(kin, level=1)
    assert bgz == kin + ".bgz" and fasta.is_bgzf(bgz)
    h = Header("p", input_file=fa, kmer_len=K)
    assert np.array_equal(h.read_table(bgz), table)
    assert np.array_equal(h.read_table(kin), table)
    assert bgzf.read_all(bgz) == table.tobytes()
    out = np.empty(table.size + 100, dtype=np.uint8)
    assert fasta.bgzf_read_into(bgz, out, threads=2) == table.size and np.array_equal(out[:table.size], table)
    with pytest.raises(ValueError):
        fasta.bgzf_read_into(bgz, np.empty(table.size - 70_000, dtype=np.uint8))
    wrong = Header("p", input_file=fa, kmer_len=K + 2)
    with pytest.raises((AssertionError, ValueError)):
        wrong.read_table(bgz)


@pytest.mark.parametrize("native", [True, False])
def test_bgzf_writer_native_and_python_paths(tmp_path, native):
    """pk_bgzf_deflate (all cores, csrc/ingest.cpp) and the zlib thread pool write files that gzip
    (the reference's reader, tools.py:296-302) gives back byte for byte; sizes around the 0xFF00
    member boundary, several batches, every level the workflow uses (bgzip -l 9, README.md:26)."""
    from pykmer_b200 import bgzf
    rng = np.random.default_rng(7)
    for n, level, batch in ((0, 6, 4), (1, 9, 4), (0xFF00 - 1, 1, 4), (0xFF00, 6, 4), (0xFF00 + 1, 9, 4),
                            (5 * 0xFF00 + 123, 9, 2), (300_000, 0, 3)):
        raw = (rng.integers(0, 256, n, dtype=np.uint8) * (rng.random(n) < 0.2)).astype(np.uint8).tobytes()
        src = str(tmp_path / f"t{n}.kin")
        open(src, "wb").write(raw)
        out = bgzf.compress_file(src, level=level, batch=batch, index=True, native=native, threads=3)
        blob = open(out, "rb").read()
        assert blob.endswith(bgzf.EOF_BLOCK) and gzip.decompress(blob) == raw
        assert bgzf.read_all(out) == raw
        # the index: one entry per data member, first implied; it agrees with the headers
        assert bgzf.read_index(out + ".gzi") == (bgzf.scan_index(out) or [(0, 0)])
        assert len(bgzf.read_index(out + ".gzi")) == max(1, -(-n // 0xFF00))


def test_bgzf_native_writer_equals_python_writer(tmp_path):
    from pykmer_b200 import bgzf
    table = synth.synth_table(4, 9)
    src = str(tmp_path / "s.09.kin")
    table.tofile(src)
    a = bgzf.compress_file(src, dst=src + ".a.bgz", level=9, native=True)
    b = bgzf.compress_file(src, dst=src + ".b.bgz", level=9, native=False)
    assert open(a, "rb").read() == open(b, "rb").read()          # same zlib, same members


def test_gzi_index_layout_and_range_reads(tmp_path, capsys):
    """.gzi layout (gzireader.py:12-19): uint64 count, then (compressed, uncompressed) offset pairs
    of every member after the first; read_range inflates only the members a slice needs."""
    import struct
    from pykmer_b200 import bgzf
    K = 9
    table = synth.synth_table(5, K)
    src = str(tmp_path / f"s.fa.{K:02d}.kin")
    table.tofile(src)
    out = bgzf.compress_file(src, level=6, index=True)
    blob = open(out + ".gzi", "rb").read()
    (count,) = struct.unpack_from("<Q", blob)
    nblk = -(-table.size // 0xFF00)
    assert count == nblk - 1 and len(blob) == 8 + 16 * count
    pairs = np.frombuffer(blob, dtype="<u8", offset=8).reshape(-1, 2)
    assert np.array_equal(pairs[:, 1], np.arange(1, nblk, dtype=np.uint64) * 0xFF00)
    comp = open(out, "rb").read()
    for c, _ in pairs:                                           # every offset is a member header
        assert comp[int(c):int(c) + 4] == b"\x1f\x8b\x08\x04"
    os.remove(out + ".gzi")
    assert bgzf.build_index(out) == out + ".gzi" and open(out + ".gzi", "rb").read() == blob
    for lo, hi in ((0, 0), (0, 1), (0, table.size), (0xFF00 - 1, 0xFF00 + 1), (0xFF00, 2 * 0xFF00),
                   (123_456, 200_000), (table.size - 5, table.size), (table.size // 4, table.size // 2)):
        assert np.array_equal(bgzf.read_range(out, lo, hi), table[lo:hi]), (lo, hi)
    os.remove(out + ".gzi")                                      # falls back to the member headers
    assert np.array_equal(bgzf.read_range(out, 70_000, 140_000), table[70_000:140_000])
    with pytest.raises(OSError):
        bgzf.read_range(out, table.size - 5, table.size + 5)
    bgzf.build_index(out)
    bgzf.print_index(out + ".gzi")
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == f"number_entries: {count:15,d}" and lines[1] == f"filesize      : {len(comp):15,d}"
    assert lines[2] == f"pos: {0:15,d} compressed_offset {int(pairs[0, 0]):15,d} uncompressed_offset {0xFF00:15,d}"
    assert len(lines) == count + 4
    with open(out + ".gzi", "wb") as fh:
        fh.write(blob[:-3])
    with pytest.raises(OSError):
        bgzf.read_index(out + ".gzi")


def test_known_answer_generator_matches_the_reference_construction(tmp_path):
    """test.py (drop-in for the reference's test.py:8-33): the K=5 file equals, text for text, the
    input the reference itself indexed for tests/golden (allkmers_05: every 5-mer as a record)."""
    res = subprocess.run([os.sys.executable, os.path.join(ROOT, "test.py"), "5", "3"], cwd=tmp_path,
                         capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.split() == ["5", "3"]
    made = gzip.open(tmp_path / "examples" / "example--05.fasta.gz").read()
    assert made == gzip.open(os.path.join(GOLD, "inputs", "allkmers_05.fasta.gz")).read()
    assert gzip.open(tmp_path / "examples" / "example--03.fasta.gz").read().count(b">") == 64


def test_merge_tables_streams_slabs_in_order(tmp_path, monkeypatch):
    """merger.merge_tables host logic (slab ring, helper-thread reads, raw / BGZF / plain-gzip
    sources) with the device layer replaced by tests/fake_device.py: several slabs per sample,
    result equal to the oracle's pair loop."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_device
    import multirank
    import pykmer_b200
    from pykmer_b200 import merger
    monkeypatch.setitem(_sys.modules, "pykmer_b200.device", fake_device)
    monkeypatch.setattr(pykmer_b200, "device", fake_device, raising=False)
    for packed in (False, True):
        sub = tmp_path / ("bgzf" if packed else "gz")
        sub.mkdir()
        kins, tables = multirank.golden_merger_inputs(sub, bgzf_packed=packed)
        headers = [Header(k, index_file=k) for k in kins]
        want = oracle.merge_matrix(np.stack(list(tables)), 2, 10)
        for slab in (4096, 16384, 1 << 20):                      # 4 slabs, exactly one, larger than the table
            got = merger.merge_tables(headers, 2, 10, slab_bytes=slab)
            assert np.array_equal(got, want), (packed, slab)


@pytest.mark.parametrize("path", ["auto", "scalar"])
def test_fasta_clean_simd_and_scalar_agree_with_python(path):
    """pk_fasta_clean (AVX-512 VBMI2 compress, or the scalar line copier: PYKMER_B200_CLEAN=scalar) against
    bytes.translate on text with every kind of line end, sizes around the 64-byte blocks and the thread
    cuts, and the two flags (other white space, non-ASCII)."""
    code = """
import ctypes, numpy as np, sys
sys.path.insert(0, %r)
from pykmer_b200 import _native as nat
rng = np.random.default_rng(3)
def clean(buf, threads):
    src = np.frombuffer(buf, dtype=np.uint8)
    dst = np.full(src.size + 80, 0xEE, dtype=np.uint8)
    kept, flags = ctypes.c_size_t(0), ctypes.c_uint32(0)
    nat.check(nat.lib.pk_fasta_clean(src.ctypes.data, src.size, dst.ctypes.data, ctypes.byref(kept), ctypes.byref(flags), threads))
    return dst, kept.value, flags.value
for n in (0, 1, 63, 64, 65, 127, 128, 1000, 4097, (1 << 20) + 77, (3 << 20) + 5):
    for style in range(4):
        body = rng.choice(np.frombuffer(b"ACGTNacgtn", dtype=np.uint8), n).copy()
        if n:
            step = (61, 71, 64, 1)[style]
            body[step - 1::step] = 10                                  # \\n
            if style == 1: body[step - 2::step] = 13                    # \\r\\n
            if style == 2 and n > 200: body[100] = 13                   # a lone \\r inside a line
        buf = body.tobytes()
        want = buf.translate(None, b"\\r\\n")
        for threads in (1, 3, 0):
            dst, kept, flags = clean(buf, threads)
            assert flags == 0 and kept == len(want) and dst[:kept].tobytes() == want, (n, style, threads)
            assert (dst[kept:] == 0xEE).all(), (n, style, threads)      # nothing written past the output
for extra, want_flags in ((b" ", 1), (b"\\t", 1), (b"\\x1c", 1), (b"\\x0b", 1), (b"\\xc3", 2), (b"\\x00", 0), (b"\\x1b", 0)):
    for pos in (0, 63, 64, 700, 4999):
        buf = bytearray(b"ACGT" * 1250); buf[pos:pos + 1] = extra
        dst, kept, flags = clean(bytes(buf), 2)
        assert flags == want_flags, (extra, pos, flags)
print("ok")
""" % ROOT
    env = dict(os.environ)
    if path != "auto":
        env["PYKMER_B200_CLEAN"] = path
    res = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stdout + res.stderr


def test_packed_table_golden_pins_the_format():
    """tests/golden/packed_table.npz (oracle/make_golden_packed.py): the library's host unpacker rebuilds the
    table from the committed packed bytes, and the oracle still packs the table to exactly those bytes."""
    from pykmer_b200 import device as dev
    g = np.load(os.path.join(GOLD, "packed_table.npz"))
    t = g["table"]
    assert np.array_equal(dev.table_unpack(g["bitmap"], g["chunk_off"], g["nz"], t.size), t)
    assert np.array_equal(oracle.unpack_table(g["bitmap"], g["chunk_off"], g["nz"], t.size), t)
    bm, off, nz = oracle.pack_table(t)
    assert np.array_equal(bm, g["bitmap"]) and np.array_equal(off, g["chunk_off"]) and np.array_equal(nz, g["nz"])
    off = g["chunk_off"].astype(np.int64)
    assert off[0] == 0 and off[2] - off[1] == 64 and off[3] == off[2]       # a full chunk: 64 units, an empty one: none
    assert int(g["bitmap"][16:32].min()) == 2 ** 64 - 1 and int(g["bitmap"][32:48].max()) == 0


def _sparse_table(rng, n, fill):
    t = rng.integers(1, 256, n, dtype=np.uint8)
    t[rng.random(n) >= fill] = 0
    return t


@pytest.mark.parametrize("isa", ["auto", "bmi2", "scalar"])
def test_table_unpack_rebuilds_packed_slices(isa, tmp_path):
    """Host half of the packed table transfer (csrc/unpack.cpp) on every code path: AVX-512 VBMI2 expand,
    BMI2 pdep, plain loop -- each in its own process, the path is chosen once (PYKMER_B200_UNPACK)."""
    code = """
import numpy as np, sys
sys.path.insert(0, %r)
from oracle import oracle
from pykmer_b200 import device as dev
rng = np.random.default_rng(5)
for n, fill in ((1024, 0.0), (1024, 1.0), (4096, 0.5), (1 << 20, 0.25), (1 << 20, 0.03), (3 << 18, 0.9)):
    t = rng.integers(1, 256, n, dtype=np.uint8); t[rng.random(n) >= fill] = 0
    bm, off, nz = oracle.pack_table(t)
    assert np.array_equal(oracle.unpack_table(bm, off, nz, n), t)
    for threads in (1, 5):
        assert np.array_equal(dev.table_unpack(bm, off, nz, n, threads=threads), t), (n, fill, threads)
    # chunks in any order: reverse them
    per = (np.unpackbits(bm.view(np.uint8), bitorder="little").reshape(-1, 1024).sum(axis=1) + 15) // 16
    roff = (np.concatenate(([0], np.cumsum(per[::-1])[:-1]))[::-1]).astype(np.uint32)
    rnz = np.concatenate([nz[int(o) * 16:(int(o) + int(p)) * 16] for o, p in zip(off[::-1], per[::-1])] + [np.zeros(0, np.uint8)])
    out = np.full(n + 64, 7, dtype=np.uint8)               # unaligned destination, guard bytes
    dev.table_unpack(bm, roff, rnz, n, out=out[3:3 + n])
    assert np.array_equal(out[3:3 + n], t) and (out[:3] == 7).all() and (out[3 + n:] == 7).all()
# a chunk that would run past the packed bytes is refused
t = np.ones(2048, dtype=np.uint8)
bm, off, nz = oracle.pack_table(t)
try:
    dev.table_unpack(bm, off, nz[:-16], 2048)
    raise SystemExit("accepted a truncated packed slice")
except ValueError:
    pass
try:
    dev.table_unpack(bm[:8], off[:1], nz, 1000)
    raise SystemExit("accepted a slice that is no whole number of chunks")
except ValueError:
    pass
print("ok")
""" % ROOT
    env = dict(os.environ)
    if isa != "auto":
        env["PYKMER_B200_UNPACK"] = isa
    res = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stdout + res.stderr


def test_library_carries_the_tensor_core_instructions():
    """The built sm_100a library really holds the 5th-generation tensor-core code paths: UTCOMMA
    (tcgen05.mma kind::mxf4, gram_f4.cu), UTCIMMA (kind::i8, gram_i8.cu), TMEM loads / stores."""
    import shutil as _shutil
    tool = _shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not installed")
    from pykmer_b200 import _native as nat
    res = subprocess.run([tool, "-sass", "-arch", "sm_100a", nat.LIB_PATH], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-500:]
    sass = res.stdout
    for mnemonic in ("UTCOMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG"):     # UTMALDG: the Gram kernel's TMA loads
        assert mnemonic in sass, mnemonic
    assert "k_gram_f4" in sass and "k_gram_i8" in sass and "k_table_pack" in sass
