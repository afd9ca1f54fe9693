"""Timing sweep of the Gram kernels over N: the integer tensor-core kernel (k_gram_i8, row-major
masks), and the FP4 kernel on tiled masks with the A operand from tensor memory (default) or from
shared memory (PYKMER_B200_GRAM_A=smem, read once per process -> one subprocess per variant), with
its diagnostic switches (PYKMER_B200_GRAM_DIAG: 1 = no global loads, 2 = no MMAs, 4 = no operand
stores).  Random masks, K=14-sized (2^23 words per sample); CUDA events, 3 warm-ups + 5 timed
launches.  Development aid -- prints one line per configuration.

    python tools/gram_sweep.py                 # both variants
    SWEEP_NS=50,255 python tools/gram_sweep.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORDS = 1 << 23


def timed(fn):
    import torch
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / 5


def child():
    import torch
    from pykmer_b200 import device as dev
    variant = os.environ.get("PYKMER_B200_GRAM_A", "tmem")
    g = torch.Generator(device="cuda").manual_seed(1)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (256, WORDS), dtype=torch.int32, device="cuda", generator=g)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    steps = WORDS / 2 / sms                                   # K=64 steps per CTA
    for n in [int(v) for v in os.environ.get("SWEEP_NS", "3,50,64,100,128,200,255").split(",")]:
        G = torch.zeros((n, n), dtype=torch.int64, device="cuda")
        rows = bits[:n].contiguous()
        os.environ["PYKMER_B200_GRAM"] = "i8"
        ref_ms = timed(lambda: dev.gram(rows, words=WORDS, out=G, accumulate=False))
        ref = G.clone()
        os.environ.pop("PYKMER_B200_GRAM")
        if variant == "tmem":
            print(f"N={n:3d} i8                 {ref_ms:7.3f} ms   {ref_ms * 1.965e6 / (2 * steps):6.1f} clk per K=32 step", flush=True)
        tiled = rows.view(n, WORDS // 32, 32).permute(1, 0, 2).contiguous().view(-1)
        for diag in [int(v) for v in os.environ.get("SWEEP_DIAGS", "0,1,2,3,4,8").split(",")]:
            os.environ["PYKMER_B200_GRAM_DIAG"] = str(diag)
            ms = timed(lambda: dev.gram_tiled(tiled, n, WORDS, out=G, accumulate=False))
            ok = "" if diag else (" exact" if torch.equal(G, ref) else " MISMATCH")
            print(f"N={n:3d} f4 A={variant:4s} diag={diag} {ms:7.3f} ms   {ms * 1.965e6 / steps:6.1f} clk per K=64 step{ok}",
                  flush=True)
        os.environ.pop("PYKMER_B200_GRAM_DIAG")
        del tiled, rows


def main():
    if os.environ.get("SWEEP_CHILD"):
        return child()
    for variant in os.environ.get("SWEEP_VARIANTS", "tmem,smem").split(","):
        env = dict(os.environ, SWEEP_CHILD="1", PYKMER_B200_GRAM_A=variant)
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)


if __name__ == "__main__":
    main()
