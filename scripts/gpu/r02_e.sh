python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py tests/test_gpu_fuzz.py tests/test_lean.py -m gpu -x -q 2>&1 | tail -6
for t in 2 3; do PLAAC_TRACKS=$t python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, plaac_b200, bench
L = plaac_b200.lib(); dev = torch.device("cuda", 0)
sc = plaac_b200.Scorer()
print("PLAAC_TRACKS", os.environ.get("PLAAC_TRACKS"), json.dumps(bench.measure_per_residue(L, sc, dev, 6547.8)))
PY
done
