python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_path_other" 2>&1 | tail -30
