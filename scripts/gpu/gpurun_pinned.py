"""plaac_score() end to end from pageable, pinned and write-combined host memory (same shard, same box)."""
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
h_codes = codes[:ntotal].cpu().numpy(); h_off = offsets.cpu().numpy()
del codes, offsets, lens
sc = plaac_b200.Scorer(device=0)
def run(tag, cptr, optr, sptr, reps=4):
    ts = []
    for it in range(reps):
        t0 = time.perf_counter(); sc.score_ptr(cptr, optr, nprot, sptr); ts.append(time.perf_counter() - t0)
    print("%-28s ms per call: %s -> %.3g residues/s" % (tag, [round(t * 1e3, 1) for t in ts], ntotal / min(ts[1:])), flush=True)
out_pg = np.zeros(nprot, dtype=plaac_b200.SUMMARY_DTYPE)
run("pageable (numpy)", h_codes.ctypes.data, h_off.ctypes.data, out_pg.ctypes.data, reps=3)
po = plaac_b200.PinnedBuffer(nprot + 1, np.int64); po.array[:] = h_off
ps = plaac_b200.PinnedBuffer(nprot, plaac_b200.SUMMARY_DTYPE)
for wc in (False, True, False, True):
    t0 = time.perf_counter()
    pc = plaac_b200.PinnedBuffer(ntotal, np.uint8, write_combined=wc)
    t1 = time.perf_counter(); pc.array[:] = h_codes; t2 = time.perf_counter()
    print("alloc %.0f ms, fill %.0f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    run("pinned, codes WC" if wc else "pinned", pc.ptr, po.ptr, ps.ptr)
    assert ps.array.tobytes() == out_pg.tobytes()
    pc.close()
t0 = time.perf_counter(); plaac_b200.host_register(h_codes); plaac_b200.host_register(h_off); plaac_b200.host_register(out_pg); t1 = time.perf_counter()
print("host_register of the three numpy arrays: %.0f ms" % ((t1 - t0) * 1e3))
run("registered numpy", h_codes.ctypes.data, h_off.ctypes.data, out_pg.ctypes.data)
