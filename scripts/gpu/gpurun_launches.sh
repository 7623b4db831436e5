ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 2 --warmup 1 --proteins-per-gpu 4000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_tmp.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(f"{k[:50]:50s} n={len(v):3d} avg={sum(v)/len(v)/1e6:9.3f} ms")
PY
