python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "v2_and_v3 or anchor" 2>&1 | tail -3
PYTHONPATH=. python - <<'PY'
import numpy as np
from tests import synth
names = "XACDEFGHIKLMNPQRSTVWY*"
lut = np.frombuffer(names.encode(), dtype=np.uint8)
codes, offs = synth.proteome(300000, seed=77)
txt = lut[codes]
chunks = []
for i in range(300000):
    s = txt[offs[i]:offs[i + 1]]
    nl = (len(s) + 59) // 60
    buf = np.full(len(s) + nl, 10, dtype=np.uint8)
    idx = np.arange(len(s))
    buf[idx + idx // 60] = s
    chunks.append(b">sp|P%07d|PROT_%d some description\n" % (i, i))
    chunks.append(buf.tobytes())
open("/tmp/big.fa", "wb").write(b"".join(chunks))
PY
for i in 1 2; do PLAAC_CLI_TIMING=1 plaac_b200/bin/plaac -i /tmp/big.fa > /tmp/out.tsv; done
PLAAC_CLI_TIMING=1 plaac_b200/bin/plaac -i /tmp/big.fa -B tests/golden/bg_freqs_HUMAN.txt > /tmp/out.tsv
