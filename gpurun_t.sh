python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "long" 2>&1 | tail -30
