python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py tests/test_gpu_fuzz.py -m gpu -q -k "residue" 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['per_residue_mode']['yeast_sized_6k']['ms'], d['per_residue_mode']['batch_200k']['ms'], d['per_residue_mode']['batch_200k']['frac_of_hbm_write_roofline'])"
