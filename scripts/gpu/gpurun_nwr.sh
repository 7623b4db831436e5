python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for n in 13 12; do
PLAAC_V2_WARP_PAIRS=$n python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-per-residue --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', 'value %.4g ms/step %.2f kernel_ms %.2f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
done
