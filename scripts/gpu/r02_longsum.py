"""Summary-mode long path: one protein alone, hybrid (k_long_post cluster + k_long_score lite + k_long_final) vs k_long_score alone."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200
from tests import synth
dev = torch.device("cuda", 0)
for n in (9000, 35000, 100000):
    codes_h, offs_h = synth.long_proteins(seed=1005, lengths=(n,))
    codes = torch.from_numpy(np.concatenate([codes_h, np.zeros(64, np.uint8)])).to(dev)
    offsets = torch.from_numpy(offs_h).to(dev)
    out = torch.zeros(160, dtype=torch.uint8, device=dev)
    res = {}
    recs = {}
    for tag, env in (("hybrid", "1"), ("k_long_score_alone", "0")):
        os.environ["PLAAC_LONG_HYBRID"] = env
        sc = plaac_b200.Scorer(); sc.set_long_path(1024)
        ms = []
        for it in range(6):
            sc.score_device(codes.data_ptr(), offsets.data_ptr(), 1, n, out.data_ptr(), sync=True)
            if it >= 2: ms.append(sc.stats().last_total_ms)
        res[tag] = sum(ms) / len(ms)
        recs[tag] = out.cpu().numpy().tobytes()
        sc.close()
    print(n, {k: round(v, 3) for k, v in res.items()}, "same bytes:", recs["hybrid"] == recs["k_long_score_alone"], flush=True)
