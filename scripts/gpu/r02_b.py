"""v3 vs v2 summary kernel on one box: records byte-identical at 4M proteins, then kernel time per variant."""
import json, os, subprocess, sys
sys.path.insert(0, os.getcwd())

def run(env, n=4000000):
    e = dict(os.environ); e.update(env)
    code = r'''
import sys, os, math, json, hashlib
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import plaac_b200, bench
L = plaac_b200.lib()
dev = torch.device("cuda", 0)
nprot = %d
lens = torch.empty(nprot, dtype=torch.int64, device=dev)
L.plaac_bench_synth_lengths(None, bench.SEED, 0, nprot, bench.LN_MEDIAN, bench.SIGMA, bench.MIN_LEN, bench.MAX_LEN, lens.data_ptr())
offsets = torch.zeros(nprot + 1, dtype=torch.int64, device=dev)
torch.cumsum(lens, 0, out=offsets[1:])
ntotal = int(offsets[-1].item())
codes = torch.empty(ntotal + 64, dtype=torch.uint8, device=dev)
bg = np.array(bench.BG_SCER, dtype=np.float64); prd = np.array(bench.PRD_28, dtype=np.float64)
L.plaac_bench_synth_residues(None, bench.SEED, 0, nprot, offsets.data_ptr(), bg.ctypes.data, prd.ctypes.data, bench.PRD_RATE, bench.X_RATE, codes.data_ptr())
out = torch.zeros(nprot * 160, dtype=torch.uint8, device=dev)
sc = plaac_b200.Scorer()
ms = []
for it in range(8):
    sc.score_device(codes.data_ptr(), offsets.data_ptr(), nprot, ntotal, out.data_ptr(), sync=True)
    if it >= 3: ms.append((sc.stats().last_score_ms, sc.stats().last_total_ms))
h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()
k = sum(m[0] for m in ms) / len(ms); t = sum(m[1] for m in ms) / len(ms)
print(json.dumps({"kernel_ms": k, "total_ms": t, "residues": ntotal, "sha": h[:16], "frac": 67.0 * ntotal / (k * 1e-3) / 18.456e12}))
''' % n
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    if r.returncode != 0:
        return {"error": r.stderr[-400:]}
    return json.loads(r.stdout.strip().splitlines()[-1])

variants = [
    ("default", {}),
    ("v3_opt1", {"PLAAC_V3_OPT": "1"}),
    ("v3_opt1_mix", {"PLAAC_V3_OPT": "1", "PLAAC_V3_MIX": "1"}),
    ("v3_onlyA", {"PLAAC_V3_OPT": "1", "PLAAC_V3_MIX": "256"}),
    ("v3_onlyB", {"PLAAC_V3_OPT": "1", "PLAAC_V3_MIX": "512"}),
    ("v3_onlyA_mix", {"PLAAC_V3_OPT": "1", "PLAAC_V3_MIX": "257"}),
    ("v3_onlyB_mix", {"PLAAC_V3_OPT": "1", "PLAAC_V3_MIX": "513"}),
    ("v3_onlyA_opt0", {"PLAAC_V3_OPT": "0", "PLAAC_V3_MIX": "256"}),
]
if len(sys.argv) > 1:
    variants = [v for v in variants if v[0] in sys.argv[1:]]
res = {}
for name, env in variants:
    res[name] = run(env)
    print(name, res[name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_b_variants.json", "w"), indent=1)
