"""Development aid (GPU box): time k_table_pack and the host unpack team apart, on a synthetic table with a
genome-like count distribution (25 % fill, geometric counts)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykmer_b200 import device as dev, _native as nat
lib = nat.lib
n = 1 << int(os.environ.get("PROBE_LOG2", 28))
rng = np.random.default_rng(1)
t = np.minimum(rng.geometric(0.5, n), 255).astype(np.uint8)
t[rng.random(n) >= float(os.environ.get("PROBE_FILL", 0.25))] = 0
d = torch.from_numpy(t).cuda()
bitmap = torch.empty(n // 64, dtype=torch.int64, device="cuda")
off = torch.empty(n // 1024, dtype=torch.int32, device="cuda")
nz = torch.empty(n + n // 64 + 16, dtype=torch.uint8, device="cuda")
units = ctypes.c_uint32(0)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nat.check(lib.pk_table_pack_device(d.data_ptr(), n, bitmap.data_ptr(), off.data_ptr(), nz.data_ptr(), ctypes.byref(units), None))
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    print(f"k_table_pack {n >> 20} Mi entries: {ms:.3f} ms  ({ms / (n >> 24) * 1e3:.1f} us per 2^24-entry window), packed {units.value * 16 / 1e6 + n / 8e6 + n / 256e6:.1f} MB")
bm, co, z = bitmap.cpu().numpy().view(np.uint64), off.cpu().numpy().view(np.uint32), nz[: units.value * 16 + 4096].cpu().numpy()
out = dev.pinned_empty(n).numpy()
for threads in (1, 4, 8, 12, 15, 16):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); dev.table_unpack(bm, co, z, n, threads=threads, out=out); best = min(best, time.perf_counter() - t0)
    print(f"pk_table_unpack (with its checks) {threads:2d} threads: {best * 1e3:.2f} ms  {n / best / 1e9:.1f} GB/s of table")
assert np.array_equal(out, t)
# plain parallel memset / copy of the same size for scale
import threading
def par(fn, threads):
    parts = np.array_split(np.arange(n // 4096), threads)
    th = [threading.Thread(target=fn, args=(int(p[0]) * 4096, (int(p[-1]) + 1) * 4096)) for p in parts]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; return time.perf_counter() - t0
for threads in (8, 15):
    s = min(par(lambda a, b: out[a:b].fill(0), threads) for _ in range(3))
    c = min(par(lambda a, b: np.copyto(out[a:b], t[a:b]), threads) for _ in range(3))
    print(f"numpy fill {threads} threads {n / s / 1e9:.1f} GB/s, copy {n / c / 1e9:.1f} GB/s")
