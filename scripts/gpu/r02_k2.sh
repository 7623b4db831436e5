mkdir -p gpurun_out
for cfg in "9000" "35000" "100000"; do
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
