"""k_pack: time outside the scoring kernel (total - kernel) at 4 M proteins per resident-CTA count, records hashed
(50cbe8baa7d723a1 = the records of the build before the full-slot fast path)."""
import importlib.util, json, os, sys
spec = importlib.util.spec_from_file_location("r02_b", os.path.join(os.path.dirname(__file__), "r02_b.py"))
src = open(spec.origin).read().split("variants = [")[0]
ns = {}
exec(compile(src, spec.origin, "exec"), ns)
res = {}
for name, env in [("default", {}), ("ctas8", {"PLAAC_PACK_CTAS": "8"}), ("ctas5", {"PLAAC_PACK_CTAS": "5"})]:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    r = ns["run"](env)
    if "error" not in r:
        r["outside_kernel_ms"] = r["total_ms"] - r["kernel_ms"]
    res[name] = r
    print(name, r, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_pack.json", "w"), indent=1)
