mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_final_tests.txt
cat gpurun_out/r02_final_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -2 gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('value %.4g ms/step %.2f kernel_ms %.2f frac %.3f launches %s'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'], d['gpu_launches']))
e=d['e2e']; print('e2e %.4g %.1f ms'%(e['value'], e['ms_per_step']), {k:round(v['value']/1e10,2) for k,v in e['variants'].items()})
print('cpu', d['cpu_baseline'])
p=d['per_residue_mode']; print({k:(round(v['ms'],3), round(v['frac_of_hbm_write_roofline'],3)) for k,v in p.items() if isinstance(v,dict)})
print(d['extras']['long_sequences'])
r=json.load(open('gpurun_out/r02_bench_ref_final.json')); print('ref', r['value'], r.get('cpu_baseline'))
PY
