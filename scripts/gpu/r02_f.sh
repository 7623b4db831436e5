for t in 2 3; do
PLAAC_TRACKS=$t ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_res_launches_t$t.csv python scripts/gpu/res_once.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r02_res_launches_t$t.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]
agg={}
for r in rows[hdr+1:]:
    d=dict(zip(H,r))
    k=d["Kernel Name"][:40]; m=d["Metric Name"]; v=float(d["Metric Value"].replace(",",""))
    agg.setdefault(k,{}).setdefault(m,[]).append(v)
print("PLAAC_TRACKS=$t (last launch of each kernel)")
for k,v in agg.items():
    t=v.get("gpu__time_duration.sum",[0])[-1]; rd=v.get("dram__bytes_read.sum",[0])[-1]; wr=v.get("dram__bytes_write.sum",[0])[-1]
    print("  %-40s %8.3f ms  read %7.1f MB  write %7.1f MB"%(k,t/1e6 if t>1e4 else t,rd/1e6 if rd>1e3 else rd, wr/1e6 if wr>1e3 else wr))
PY
done
