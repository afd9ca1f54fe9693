"""Timing sweep of the Gram kernels (i8 / f4) over N, the f4 producer variants (PYKMER_B200_GRAM_AHEAD: groups x depth) and its diagnostic switches.
Random masks, K=14-sized (2^23 words per sample); CUDA events, 3 warm-ups + 5 timed launches.
Development aid -- prints one line per configuration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykmer_b200 import device as dev  # noqa: E402

WORDS = 1 << 23
PAD = int(os.environ.get("SWEEP_PAD_WORDS", "0"))       # extra words between rows (row stride = WORDS + PAD)


def timed(bits, n, env):
    for k in ("PYKMER_B200_GRAM", "PYKMER_B200_GRAM_AHEAD", "PYKMER_B200_GRAM_DIAG", "PYKMER_B200_GRAM_TILED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    G = torch.zeros((n, n), dtype=torch.int64, device="cuda")
    for _ in range(3):
        dev.gram(bits[:n], words=WORDS, out=G, accumulate=False) if PAD == 0 else gram_strided(bits, n, G)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        dev.gram(bits[:n], words=WORDS, out=G, accumulate=False) if PAD == 0 else gram_strided(bits, n, G)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / 5, G


def gram_strided(bits, n, G):
    """dev.gram insists on a contiguous 2-D tensor; a padded row stride goes straight to the C ABI."""
    from pykmer_b200 import _native as nat
    nat.check(nat.lib.pk_gram_device(bits.data_ptr(), n, WORDS, WORDS + PAD, G.data_ptr(), 0,
                                     torch.cuda.current_stream().cuda_stream))
    return G


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (256, WORDS + PAD), dtype=torch.int32, device="cuda", generator=g)
    print(f"row stride = {WORDS} + {PAD} words", flush=True)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for n in [int(v) for v in os.environ.get("SWEEP_NS", "50,64,100,128,200,255").split(",")]:
        ref_ms, ref = timed(bits, n, {"PYKMER_B200_GRAM": "i8"})
        steps = WORDS / 2 / sms
        print(f"N={n:3d} i8            {ref_ms:7.3f} ms   {ref_ms * 1.965e6 / (2 * steps):6.1f} clk per K=32 step", flush=True)
        # the same masks in the tiled layout: [WORDS / 32][n rows][32 words]
        tiled = bits[:n, :WORDS].reshape(n, WORDS // 32, 32).permute(1, 0, 2).contiguous().reshape(n, WORDS)
        for layout, src in (("rows ", bits), ("tiled", tiled)):
            for diag in (0, 2):
                ms, G = timed(src, n, {"PYKMER_B200_GRAM": "f4", "PYKMER_B200_GRAM_DIAG": str(diag),
                                       "PYKMER_B200_GRAM_TILED": "0" if layout == "rows " else "1"})
                ok = "" if diag else (" exact" if torch.equal(G, ref) else " MISMATCH")
                print(f"N={n:3d} f4 {layout} diag={diag} {ms:7.3f} ms   {ms * 1.965e6 / steps:6.1f} clk per K=64 step{ok}",
                      flush=True)
        del tiled


if __name__ == "__main__":
    main()
