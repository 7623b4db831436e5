mkdir -p gpurun_out
for cfg in "9000" "35000 10000000" "35000 1024" "100000 10000000" "100000 1024"; do
  tag=$(echo $cfg | tr ' ' '_')
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_lr_$tag.csv python scripts/gpu/r02_longres_one.py $cfg > gpurun_out/r02_lr_$tag.log 2>&1
  tail -1 gpurun_out/r02_lr_$tag.log
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r02_lr_$tag.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; last={}
for r in rows[hdr+1:]:
    d=dict(zip(H,r)); last[d["Kernel Name"][:34]]=float(d["Metric Value"].replace(",",""))
print("$tag", {k: round(v/1e3,1) for k,v in last.items() if v>8000})
PY
done
cat > /tmp/six.py <<'PY'
import os, sys, math
sys.path.insert(0, os.getcwd())
import numpy as np, torch, plaac_b200, bench
L = plaac_b200.lib(); dev = torch.device("cuda", 0)
nprot = 6000
lens = torch.empty(nprot, dtype=torch.int64, device=dev)
L.plaac_bench_synth_lengths(None, 1001, 0, nprot, math.log(407.0), 0.66, bench.MIN_LEN, bench.MAX_LEN, lens.data_ptr())
offsets = torch.zeros(nprot + 1, dtype=torch.int64, device=dev)
torch.cumsum(lens, 0, out=offsets[1:])
ntotal = int(offsets[-1].item())
codes = torch.empty(ntotal + 64, dtype=torch.uint8, device=dev)
bg = np.array(bench.BG_SCER, dtype=np.float64); prd = np.array(bench.PRD_28, dtype=np.float64)
L.plaac_bench_synth_residues(None, 1001, 0, nprot, offsets.data_ptr(), bg.ctypes.data, prd.ctypes.data, bench.PRD_RATE, bench.X_RATE, codes.data_ptr())
u8 = torch.empty(2 * ntotal, dtype=torch.uint8, device=dev)
stride = (ntotal + 3) & ~3
f64 = torch.empty(10 * stride, dtype=torch.float64, device=dev)
ptrs = {"vit": u8.data_ptr(), "map": u8.data_ptr() + ntotal}
for k, nm in enumerate(plaac_b200.RESIDUE_F64): ptrs[nm] = f64.data_ptr() + 8 * k * stride
sc = plaac_b200.Scorer(); sc.set_long_path(-1)
for _ in range(3):
    sc.score_device(codes.data_ptr(), offsets.data_ptr(), nprot, ntotal, 0, residue_ptrs=ptrs, sync=True)
print("6k ms", sc.stats().last_total_ms)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_lr_6k.csv python /tmp/six.py > gpurun_out/r02_lr_6k.log 2>&1
tail -1 gpurun_out/r02_lr_6k.log
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r02_lr_6k.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; last={}
for r in rows[hdr+1:]:
    d=dict(zip(H,r)); last[d["Kernel Name"][:34]]=float(d["Metric Value"].replace(",",""))
print("6k", {k: round(v/1e3,1) for k,v in last.items()})
PY
