"""Launch the drop-in CLIs as a world_size-N torch.distributed job on this machine (gloo
rendezvous on 127.0.0.1) -- shared by the CPU tests (device layer replaced by tests/fake_device.py)
and the GPU tests (the real library, every rank on cuda:0)."""
import gzip
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
fake, tool, argv = sys.argv[1] == "fake", sys.argv[2], sys.argv[3:]
if fake:
    import pykmer_b200
    import fake_device
    sys.modules["pykmer_b200.device"] = fake_device
    pykmer_b200.device = fake_device
from pykmer_b200 import indexer, merger
{"indexer": indexer, "merger": merger}[tool].main(argv)
import torch.distributed as dist
if dist.is_initialized():
    dist.destroy_process_group()
"""


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_cli(tmp_path, tool, argv, nranks=2, fake=True, timeout=300):
    """-> [(returncode, output)] per rank."""
    script = tmp_path / "rank_worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = free_port()
    procs = []
    for rank in range(nranks):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(nranks), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PYKMER_B200_DIST_BACKEND="gloo")
        procs.append(subprocess.Popen([sys.executable, str(script), "fake" if fake else "real", tool] +
                                      [str(a) for a in argv], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    res = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
            out += "\n[timed out]"
        res.append((p.returncode, out))
    return res


def golden_merger_inputs(tmp_path, bgzf_packed=False):
    """The K=7 samples of the golden merger run as .kin / .kin.bgz files with their JSONs, in the
    golden run's directory layout (so that sorted paths give its order)."""
    from pykmer_b200.tools import Header
    samples = np.load(os.path.join(GOLD, "merger", "samples_K07.npz"))
    template = json.load(open(os.path.join(GOLD, "indexer", "tiny_mixed.fa.07.json")))
    kins = []
    for name, table in zip(samples["names"], samples["tables"]):
        name = str(name)
        packed = name.endswith(".bgz")
        sub = tmp_path / "merge" if name.startswith("synth") else tmp_path
        sub.mkdir(exist_ok=True)
        kin = str(sub / (name[:-4] if packed else name))
        base = kin[:-len(".07.kin")]
        open(base, "w").close()
        meta = {k: template.get(k) for k in Header.HEADER_FIXED + Header.HEADER_DATA}
        meta.update(input_file_path=base, input_file_name=os.path.basename(base), kmer_len=7)
        json.dump(meta, open(kin + ".json", "w"))
        if packed and bgzf_packed:                      # real BGZF: slices are read member-wise
            from pykmer_b200 import bgzf
            table.tofile(kin)
            bgzf.compress_file(kin, level=6, index=True)
            os.remove(kin)
            kins.append(kin + ".bgz")
        elif packed:                                    # plain gzip under the .bgz name
            with gzip.open(kin + ".bgz", "wb") as fz:
                fz.write(table.tobytes())
            kins.append(kin + ".bgz")
        else:
            table.tofile(kin)
            kins.append(kin)
    return kins, samples["tables"]
