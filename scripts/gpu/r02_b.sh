python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -5
python scripts/gpu/r02_b.py
