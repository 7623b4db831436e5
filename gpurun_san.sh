export PATH=/usr/local/cuda/bin:$PATH
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_rank.py tests/test_gpu_parity.py -m gpu -x -q -k "rank or long_sequences or edge_cases or classic or empty" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_rank.py tests/test_gpu_parity.py -m gpu -x -q -k "rank_ties or rank_empty or long_sequences" 2>&1 | tail -15
echo "racecheck rc=$?"
