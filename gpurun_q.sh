python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-per-residue --no-e2e --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.2f kernel_ms %.2f frac %.3f launches %d'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['gpu_launches']))"; done
