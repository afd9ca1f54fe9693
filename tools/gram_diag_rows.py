import os, sys, torch
sys.path.insert(0, "/root/repo")
from pykmer_b200 import device as dev
g = torch.Generator(device="cuda").manual_seed(1)
for n in (50, 3, 100):
    words = 1 << 19
    rows = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, words), dtype=torch.int32, device="cuda", generator=g)
    os.environ["PYKMER_B200_GRAM"] = "i8"; ref = dev.gram(rows, words=words); os.environ.pop("PYKMER_B200_GRAM")
    tiled = rows.view(n, words // 32, 32).permute(1, 0, 2).contiguous().view(-1)
    for rep in range(2):
        G = dev.gram_tiled(tiled, n, words); d = (G - ref)
        badr = [(i, int((d[i] != 0).sum())) for i in range(n) if (d[i] != 0).any()]
        badc = [(j, int((d[:, j] != 0).sum())) for j in range(n) if (d[:, j] != 0).any()]
        print(f"N={n} rep={rep}: {int((d != 0).sum())} cells differ, max {int(d.abs().max())}; rows(count) {badr[:40]}; cols {badc[:12]}", flush=True)
