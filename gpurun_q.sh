python gpurun_dbg.py
python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_jar_vectors.py -m gpu -q -k "random_case_against or long" 2>&1 | tail -8
PLAAC_LONG_TIES=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_sequences or mixed_batch" 2>&1 | tail -3
python gpurun_long_time.py 2>&1 | head -4
