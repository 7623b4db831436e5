# round 2, first GPU call: all GPU tests, the bench line (new e2e transports), raw link probe at one rank
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_a_tests.txt
cat gpurun_out/r02_a_tests.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err
tail -3 gpurun_out/r02_a_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_a_bench.json'))
print('value %.4g ms/step %.2f kernel_ms %.2f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
e=d['e2e']
print('e2e', e['value'], e['ms_per_step'], e.get('hits'))
for k,v in e['variants'].items(): print(k, v['value'], v['ms_per_step'], v.get('matches_device_path'))
print(d['per_residue_mode'])
print(d['extras']['long_sequences'])
PY
python scripts/gpu/link_probe.py --out gpurun_out/r02_link_probe_1.json
