import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200
from tests import synth
n = int(sys.argv[1])
dev = torch.device("cuda", 0)
codes_h, offs_h = synth.long_proteins(seed=1005, lengths=(n,))
codes = torch.from_numpy(np.concatenate([codes_h, np.zeros(64, np.uint8)])).to(dev)
offsets = torch.from_numpy(offs_h).to(dev)
out = torch.zeros(160, dtype=torch.uint8, device=dev)
sc = plaac_b200.Scorer(); sc.set_long_path(1024)
for it in range(3):
    sc.score_device(codes.data_ptr(), offsets.data_ptr(), 1, n, out.data_ptr(), sync=True)
print(n, sc.stats().last_total_ms)
