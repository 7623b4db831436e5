PLAAC_LONG_CLOCKS=1 python - <<'PY'
import numpy as np, torch
import plaac_b200
from tests import synth
for lengths in [(100000,), (35000,), (9000,)]:
    codes, offs = synth.long_proteins(lengths=lengths)
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    d_out = torch.zeros((1, 160), dtype=torch.uint8, device="cuda")
    sc = plaac_b200.Scorer(device=0)
    for it in range(2):
        sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), 1, int(offs[-1]), d_out.data_ptr())
        st = sc.stats()
    print(lengths, st.last_total_ms, st.last_score_ms, flush=True)
    sc.close()
PY
