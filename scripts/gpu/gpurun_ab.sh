cat > /tmp/ab.py <<'PY'
import numpy as np, torch
import plaac_b200
from tests import synth
for lengths in [(100000,), (80000,), (65000,), (50000,), (35000,), (9000,)]:
    codes, offs = synth.long_proteins(lengths=lengths)
    n = len(offs) - 1
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    d_out = torch.zeros((n, 160), dtype=torch.uint8, device="cuda")
    sc = plaac_b200.Scorer(device=0)
    ts = []
    for it in range(8):
        sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), n, int(offs[-1]), d_out.data_ptr())
        st = sc.stats()
        ts.append((round(st.last_total_ms, 4), round(st.last_score_ms, 4)))
    print(lengths, min(ts), sorted(ts)[len(ts) // 2], flush=True)
    sc.close()
PY
for v in 49152; do echo "== cm_min $v"; PLAAC_LONG_CM_MIN=$v PYTHONPATH=. python /tmp/ab.py; done
python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py tests/test_gpu_fuzz.py -m gpu -x -q -k "long or random_case_against or tie" 2>&1 | tail -4
PLAAC_LONG_CM_MIN=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "long or random_case_against or tie" 2>&1 | tail -4
PLAAC_LONG_TIES=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_sequences or mixed_batch" 2>&1 | tail -2
