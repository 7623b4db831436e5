mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_m_tests.txt
cat gpurun_out/r02_m_tests.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err
tail -3 gpurun_out/r02_m_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_m_bench.json'))
print('value %.4g ms/step %.2f kernel_ms %.2f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
e=d['e2e']
print('e2e', e['value'], e['ms_per_step'])
print(json.dumps(d['per_residue_mode'], indent=1))
print(json.dumps(d['extras']['long_sequences'], indent=1))
PY
