# e2e of plaac_score() with two builds of the library on the same box (ab/libF.so = commit e9446eb, ab/libH.so = HEAD)
cat > /tmp/e2e.py <<'PY'
import time, numpy as np, torch
import plaac_b200, bench
L = plaac_b200.lib(); dev = torch.device("cuda", 0)
nprot = 12_500_000
lens = torch.empty(nprot, dtype=torch.int64, device=dev)
L.plaac_bench_synth_lengths(None, bench.SEED, 0, nprot, bench.LN_MEDIAN, bench.SIGMA, bench.MIN_LEN, bench.MAX_LEN, lens.data_ptr())
offsets = torch.zeros(nprot + 1, dtype=torch.int64, device=dev); torch.cumsum(lens, 0, out=offsets[1:])
ntotal = int(offsets[-1].item())
codes = torch.empty(ntotal + 64, dtype=torch.uint8, device=dev)
bg = np.array(bench.BG_SCER, dtype=np.float64); prd = np.array(bench.PRD_28, dtype=np.float64)
L.plaac_bench_synth_residues(None, bench.SEED, 0, nprot, offsets.data_ptr(), bg.ctypes.data, prd.ctypes.data, bench.PRD_RATE, bench.X_RATE, codes.data_ptr())
h_codes = torch.empty(ntotal, dtype=torch.uint8).pin_memory(); h_codes.copy_(codes[:ntotal])
h_off = torch.empty(nprot + 1, dtype=torch.int64).pin_memory(); h_off.copy_(offsets)
h_sum = torch.empty(nprot * 160, dtype=torch.uint8).pin_memory()
del codes, offsets, lens
torch.cuda.synchronize()
sc = plaac_b200.Scorer(device=0)
ts = []
for it in range(7):
    t0 = time.perf_counter(); sc.score_ptr(h_codes.data_ptr(), h_off.data_ptr(), nprot, h_sum.data_ptr()); ts.append(time.perf_counter() - t0)
print("e2e ms per call:", [round(t * 1e3, 1) for t in ts], flush=True)
d = torch.empty(ntotal, dtype=torch.uint8, device=dev); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(h_codes, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain H2D of the codes: %.1f ms = %.1f GB/s" % (dt * 1e3, ntotal / dt / 1e9))
dd = torch.empty(nprot * 160, dtype=torch.uint8, device=dev); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); h_sum.copy_(dd, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain D2H of the records: %.1f ms = %.1f GB/s" % (dt * 1e3, nprot * 160 / dt / 1e9))
PY
for v in F H F H; do cp ab/lib$v.so plaac_b200/libplaac_cuda.so; echo "== $v"; PYTHONPATH=. python /tmp/e2e.py; done
cp ab/libH.so plaac_b200/libplaac_cuda.so
