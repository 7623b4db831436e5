"""One long protein in per-residue mode, a few calls (for an ncu launch list).  argv: n [big_min]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200, bench
n = int(sys.argv[1])
if len(sys.argv) > 2: os.environ["PLAAC_LP_BIG_MIN"] = sys.argv[2]
dev = torch.device("cuda", 0)
sc = plaac_b200.Scorer(); sc.set_long_path(1024)
rng = np.random.default_rng(1005)
s = rng.choice(22, size=n, p=np.array(bench.BG_SCER) / np.sum(bench.BG_SCER)).astype(np.uint8)
codes = torch.from_numpy(np.concatenate([s, np.zeros(64, np.uint8)])).to(dev)
offsets = torch.tensor([0, n], dtype=torch.int64, device=dev)
u8 = torch.empty(2 * n, dtype=torch.uint8, device=dev); f64 = torch.empty(10 * n, dtype=torch.float64, device=dev)
ptrs = {"vit": u8.data_ptr(), "map": u8.data_ptr() + n}
for k, nm in enumerate(plaac_b200.RESIDUE_F64): ptrs[nm] = f64.data_ptr() + 8 * k * n
for _ in range(3):
    sc.score_device(codes.data_ptr(), offsets.data_ptr(), 1, n, 0, residue_ptrs=ptrs, sync=True)
print(n, sc.stats().last_total_ms, "redone", sc.stats().long_redone_chunks)
