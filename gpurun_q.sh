python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py tests/test_sharding.py -m gpu -x -q -k "per_residue or multi_ctx" 2>&1 | tail -3
python gpurun_res_prof.py
