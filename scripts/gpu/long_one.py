"""One long protein through the long-sequence path (for ncu): python scripts/gpu/long_one.py <residues>"""
import sys
import torch
import plaac_b200
from tests import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
codes, offs = synth.long_proteins(lengths=(n,))
d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
d_out = torch.zeros((1, 160), dtype=torch.uint8, device="cuda")
sc = plaac_b200.Scorer(device=0)
for it in range(3):
    sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), 1, int(offs[-1]), d_out.data_ptr())
print(n, sc.stats().last_total_ms)
sc.close()
