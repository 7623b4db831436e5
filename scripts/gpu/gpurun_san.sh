export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_jar_vectors.py tests/test_rank.py -m gpu -x -q -k "random_case or long_path or long_sequences or per_residue_against_golden or cuda_summary or rank_ties or edge_cases" 2>&1 | tail -6
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_sequences or tie_constants or per_residue_against_golden" 2>&1 | tail -5
echo "synccheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tie_constants or per_residue_against_golden" 2>&1 | tail -5
echo "racecheck rc=$?"
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_rank.py -m gpu -x -q -k "tie_constants or rank_ties or classic" 2>&1 | tail -8
echo "initcheck rc=$?"
