export PATH=/usr/local/cuda/bin:$PATH
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_scratch_layouts or long_sequences or chunked_host_path" 2>&1 | tail -6
echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_jar_vectors.py -m gpu -x -q -k "both_scratch_layouts or classic or edge or cuda_summary or chunked_host_path" 2>&1 | tail -8
echo "initcheck rc=$?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_scratch_layouts" 2>&1 | tail -5
echo "racecheck rc=$?"
