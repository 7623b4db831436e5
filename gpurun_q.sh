python -m pytest tests/test_gpu_parity.py tests/test_sharding.py tests/test_host_cli.py -m gpu -x -q 2>&1 | tail -3
python gpurun_e2e.py 2>&1 | tail -6
