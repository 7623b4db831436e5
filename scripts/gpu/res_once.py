import os, sys, math
sys.path.insert(0, os.getcwd())
import numpy as np, torch, plaac_b200, bench
L = plaac_b200.lib(); dev = torch.device("cuda", 0)
nprot = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
lens = torch.empty(nprot, dtype=torch.int64, device=dev)
L.plaac_bench_synth_lengths(None, 1001, 0, nprot, math.log(407.0), 0.66, bench.MIN_LEN, bench.MAX_LEN, lens.data_ptr())
offsets = torch.zeros(nprot + 1, dtype=torch.int64, device=dev)
torch.cumsum(lens, 0, out=offsets[1:])
ntotal = int(offsets[-1].item())
codes = torch.empty(ntotal + 64, dtype=torch.uint8, device=dev)
bg = np.array(bench.BG_SCER, dtype=np.float64); prd = np.array(bench.PRD_28, dtype=np.float64)
L.plaac_bench_synth_residues(None, 1001, 0, nprot, offsets.data_ptr(), bg.ctypes.data, prd.ctypes.data, bench.PRD_RATE, bench.X_RATE, codes.data_ptr())
u8 = torch.empty(2 * ntotal, dtype=torch.uint8, device=dev)
stride = (ntotal + 3) & ~3
f64 = torch.empty(10 * stride, dtype=torch.float64, device=dev)
ptrs = {"vit": u8.data_ptr(), "map": u8.data_ptr() + ntotal}
for k, nm in enumerate(plaac_b200.RESIDUE_F64):
    ptrs[nm] = f64.data_ptr() + 8 * k * stride
sc = plaac_b200.Scorer()
for _ in range(int(os.environ.get("REPS", "2"))):
    sc.score_device(codes.data_ptr(), offsets.data_ptr(), nprot, ntotal, 0, residue_ptrs=ptrs, sync=True)
print("ms", sc.stats().last_total_ms, "residues", ntotal)
