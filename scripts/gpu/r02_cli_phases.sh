# phase times of the host CLI (PLAAC_CLI_TIMING=1), 167 MB and 1.5 GB inputs
PYTHONPATH=. python - <<'PY'
import numpy as np, os
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
blob = b"".join(chunks)
open("/tmp/big.fa", "wb").write(blob)
with open("/tmp/huge.fa", "wb") as f:
    for r in range(9): f.write(blob)
print("residues", int(offs[-1]), os.path.getsize("/tmp/big.fa"), os.path.getsize("/tmp/huge.fa"))
PY
for f in /tmp/big.fa /tmp/huge.fa; do
  for args in "" "--batch-mb 64" "--batch-mb 32"; do
    for it in 1 2; do
      env PLAAC_CLI_TIMING=1 plaac_b200/bin/plaac -i $f $args > /tmp/out.tsv 2> /tmp/err.txt
    done
    echo "== $f $args"; grep -v "^$" /tmp/err.txt | tail -40
  done
done
