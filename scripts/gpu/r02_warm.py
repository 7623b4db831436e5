"""Long per-residue path: time and sequentially redone chunks against the warm-up length."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200, bench
from tests import synth
dev = torch.device("cuda", 0)
sc = plaac_b200.Scorer()
for n in (9000, 35000, 100000):
    codes_h, offs_h = synth.long_proteins(seed=1005, lengths=(n,))
    codes = torch.from_numpy(np.concatenate([codes_h, np.zeros(64, np.uint8)])).to(dev)
    offsets = torch.from_numpy(offs_h).to(dev)
    u8 = torch.empty(2 * n, dtype=torch.uint8, device=dev); f64 = torch.empty(10 * n, dtype=torch.float64, device=dev)
    ptrs = {"vit": u8.data_ptr(), "map": u8.data_ptr() + n}
    for k, nm in enumerate(plaac_b200.RESIDUE_F64): ptrs[nm] = f64.data_ptr() + 8 * k * n
    for warm in (16, 32, 64, 96, 128, 192, 256):
        sc.set_long_path(1024, warm)
        ms = []
        r0 = sc.stats().long_redone_chunks
        for it in range(5):
            sc.score_device(codes.data_ptr(), offsets.data_ptr(), 1, n, 0, residue_ptrs=ptrs, sync=True)
            if it >= 2: ms.append(sc.stats().last_total_ms)
        r1 = sc.stats().long_redone_chunks
        print(n, "warm", warm, "ms %.3f" % (sum(ms) / len(ms)), "redone per call", (r1 - r0) / 5, flush=True)
