export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python scripts/gpu/r02_san_job.py > gpurun_out/r02_san_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02_san_$tool.txt
done
