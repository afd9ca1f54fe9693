"""Development aid: where does the FP4 Gram kernel differ from the integer kernel?  Prints, per
(words, variant), the number of differing cells and the largest difference in each of the blocks
(lo,lo) (lo,hi) (hi,hi)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykmer_b200 import device as dev  # noqa: E402


def main():
    n = int(os.environ.get("DIAG_N", "255"))
    g = torch.Generator(device="cuda").manual_seed(1)
    for logw in [int(v) for v in os.environ.get("DIAG_LOGW", "15,19,23").split(",")]:
        words = 1 << logw
        rows = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, words), dtype=torch.int32, device="cuda", generator=g)
        os.environ["PYKMER_B200_GRAM"] = "i8"
        ref = dev.gram(rows, words=words)
        os.environ.pop("PYKMER_B200_GRAM")
        tiled = rows.view(n, words // 32, 32).permute(1, 0, 2).contiguous().view(-1)
        for diag in os.environ.get("DIAG_BITS", "0,8").split(","):
            os.environ["PYKMER_B200_GRAM_DIAG"] = diag
            for rep in range(int(os.environ.get("DIAG_REPS", "6"))):
                G = dev.gram_tiled(tiled, n, words)
                d = (G - ref)
                h = min(n, 128)
                blocks = {"lolo": d[:h, :h], "lohi": d[:h, h:], "hihi": d[h:, h:], "lolo_row0": d[:1, :h],
                          "lolo_col0": d[1:h, :1]}
                msg = ", ".join(f"{k}: {int((v != 0).sum())} cells, max {int(v.abs().max()) if v.numel() else 0}, "
                                f"sum {int(v.sum())}" for k, v in blocks.items())
                bad_rows = torch.nonzero((d != 0).sum(dim=1) > n // 4).flatten().tolist()
                bad_cols = torch.nonzero((d != 0).sum(dim=0) > n // 4).flatten().tolist()
                if os.environ.get("DIAG_SHORT"):
                    msg = ""
                print(f"words=2^{logw} diag={diag} rep={rep}: {msg}; rows {bad_rows[:8]} cols {bad_cols[:8]}", flush=True)
        os.environ.pop("PYKMER_B200_GRAM_DIAG")
        del rows, tiled


if __name__ == "__main__":
    main()
