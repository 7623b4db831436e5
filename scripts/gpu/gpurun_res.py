import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200, bench
dev = torch.device("cuda", 0)
L = plaac_b200.lib()
def make(nprot, median=407.0, sigma=0.66, seed=1001):
    lens = torch.empty(nprot, dtype=torch.int64, device=dev)
    L.plaac_bench_synth_lengths(None, seed, 0, nprot, float(np.log(median)), sigma, 16, 40000, lens.data_ptr())
    offsets = torch.zeros(nprot + 1, dtype=torch.int64, device=dev); torch.cumsum(lens, 0, out=offsets[1:])
    ntotal = int(offsets[-1])
    codes = torch.empty(ntotal + 64, dtype=torch.uint8, device=dev)
    bg = np.array(bench.BG_SCER); prd = np.array(bench.PRD_28)
    L.plaac_bench_synth_residues(None, seed, 0, nprot, offsets.data_ptr(), bg.ctypes.data, prd.ctypes.data, 0.05, 1e-4, codes.data_ptr())
    return codes, offsets, ntotal
sc = plaac_b200.Scorer()
for nprot in (6000, 200000):
    codes, offsets, ntotal = make(nprot)
    u8 = torch.empty(2 * ntotal, dtype=torch.uint8, device=dev)
    f64 = torch.empty(10 * ntotal, dtype=torch.float64, device=dev)
    ptrs = {"vit": u8.data_ptr(), "map": u8.data_ptr() + ntotal}
    for k, nm in enumerate(plaac_b200.RESIDUE_F64):
        ptrs[nm] = f64.data_ptr() + 8 * k * ntotal
    summ = torch.empty(nprot * 160, dtype=torch.uint8, device=dev)
    def f(): sc.score_device(codes.data_ptr(), offsets.data_ptr(), nprot, ntotal, 0, residue_ptrs=ptrs, sync=True)
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"per-residue nprot={nprot} residues={ntotal}: {dt*1e3:.2f} ms  {ntotal/dt:.3e} aa/s  out {82*ntotal/dt/1e9:.0f} GB/s (roofline {6540/83*1e9:.3e} aa/s)")
