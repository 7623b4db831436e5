python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py -m gpu -x -q 2>&1 | tail -5
python gpurun_long_time.py 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-per-residue | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.2f kernel_ms %.2f frac %.3f launches %d'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['gpu_launches'])); print(d['extras'])"
