import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200
from tests import synth
dev = torch.device("cuda", 0)
# raw PCIe
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print("H2D GB/s", n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9)
print("D2H GB/s", n / t(lambda: h2.copy_(d2, non_blocking=True)) / 1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("both: each GB/s", n / t(both) / 1e9)
del h, d, h2, d2
L = plaac_b200.lib()
nprot = 4_000_000
import bench
lens = torch.empty(nprot, dtype=torch.int64, device=dev)
L.plaac_bench_synth_lengths(None, 1004, 0, nprot, bench.LN_MEDIAN, bench.SIGMA, 16, 40000, lens.data_ptr())
offsets = torch.zeros(nprot + 1, dtype=torch.int64, device=dev); torch.cumsum(lens, 0, out=offsets[1:])
ntotal = int(offsets[-1])
codes = torch.empty(ntotal + 64, dtype=torch.uint8, device=dev)
bg = np.array(bench.BG_SCER); prd = np.array(bench.PRD_28)
L.plaac_bench_synth_residues(None, 1004, 0, nprot, offsets.data_ptr(), bg.ctypes.data, prd.ctypes.data, 0.05, 1e-4, codes.data_ptr())
hc = torch.empty(ntotal, dtype=torch.uint8).pin_memory(); hc.copy_(codes[:ntotal])
ho = torch.empty(nprot + 1, dtype=torch.int64).pin_memory(); ho.copy_(offsets)
hs = torch.empty(nprot * 160, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
sc = plaac_b200.Scorer()
for chunk in (0, 32 << 20, 64 << 20, 128 << 20, 512 << 20):
    if chunk: sc.set_chunk(max_residues=chunk)
    f = lambda: sc.score_ptr(hc.data_ptr(), ho.data_ptr(), nprot, hs.data_ptr())
    f(); f()
    dt = t(f, 3)
    print(f"chunk {chunk>>20 or 256}M: {dt*1e3:.1f} ms  {ntotal/dt:.3e} res/s  ideal H2D-only {(ntotal+8*nprot)/55e9*1e3:.1f} ms")
