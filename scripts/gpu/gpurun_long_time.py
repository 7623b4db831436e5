import numpy as np, torch, time
import plaac_b200
from tests import synth
for lengths in [(35000,), (100000,), (35000, 100000), (4910,)]:
    codes, offs = synth.long_proteins(lengths=lengths)
    n = len(offs) - 1
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    d_out = torch.zeros((n, 160), dtype=torch.uint8, device="cuda")
    for min_len in (0, 4096):
        sc = plaac_b200.Scorer(device=0); sc.set_long_path(min_len)
        ts = []
        for it in range(5):
            sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), n, int(offs[-1]), d_out.data_ptr())
            ts.append(sc.stats().last_total_ms)
        print(lengths, "long_min", min_len, "ms", [round(t, 3) for t in ts])
        sc.close()
# yeast-sized proteome latency with different thresholds
codes, offs = synth.proteome(6000, 1001)
n = len(offs) - 1
d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
d_out = torch.zeros((n, 160), dtype=torch.uint8, device="cuda")
print("yeast-sized: max len", int(np.diff(offs).max()), "residues", int(offs[-1]))
for min_len in (0, 4096, 2048, 1024):
    sc = plaac_b200.Scorer(device=0); sc.set_long_path(min_len)
    ts = []
    for it in range(5):
        sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), n, int(offs[-1]), d_out.data_ptr())
        ts.append(sc.stats().last_total_ms)
    print("yeast long_min", min_len, "ms", [round(t, 3) for t in ts], "long proteins", sc.stats().long_proteins // 5)
    sc.close()
