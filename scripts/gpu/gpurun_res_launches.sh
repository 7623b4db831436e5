PYTHONPATH=. ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/res_launches.csv python scripts/gpu/gpurun_res_prof.py > gpurun_out/res_prof.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/res_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ki][:48],r[mi],r[ui]),[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(f"{k[0]:48s} {k[1]:26s} {k[2]:6s} n={len(v):3d} avg={sum(v)/len(v):14.3f}")
PY
tail -3 gpurun_out/res_prof.log
