# wall time of the host CLI on a synthetic FASTA (300 k proteins, ~110 MB), host reader and --gpu-ingest
PYTHONPATH=. python - <<'PY'
import numpy as np
from tests import synth
codes, offs = synth.proteome(300000, seed=77)
names = "XACDEFGHIKLMNPQRSTVWY*"
lut = np.frombuffer(names.encode(), dtype=np.uint8)
txt = lut[codes]
with open("/tmp/big.fa", "wb") as f:
    for i in range(len(offs) - 1):
        f.write(b">sp|P%06d|PROT_%d some description\n" % (i, i))
        s = txt[offs[i]:offs[i + 1]].tobytes()
        for j in range(0, len(s), 60):
            f.write(s[j:j + 60] + b"\n")
print("residues", int(offs[-1]))
PY
ls -la /tmp/big.fa
python - <<'PY'
import subprocess, time
for args in ([], ["--gpu-ingest"], ["--rank-core"]):
    t0 = time.perf_counter()
    with open("/tmp/out.tsv", "wb") as f:
        subprocess.run(["plaac_b200/bin/plaac", "-i", "/tmp/big.fa", *args], stdout=f, check=True)
    dt = time.perf_counter() - t0
    n = sum(1 for _ in open("/tmp/out.tsv", "rb"))
    print(args, "%.2f s wall, %d lines" % (dt, n), flush=True)
PY
