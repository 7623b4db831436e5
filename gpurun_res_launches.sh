ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_res.csv python gpurun_res.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_res.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items():
    h=len(v)//2
    print(f"{k[:50]:50s} n={len(v):3d} small={sum(v[:h])/max(h,1)/1e6:8.3f} ms large={sum(v[h:])/max(len(v)-h,1)/1e6:8.3f} ms")
PY
