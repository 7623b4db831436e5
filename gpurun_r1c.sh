set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 600 gpurun_out/bench_c.err
ncu --set full --clock-control none --import-source on -k regex:k_score_summary_v2 -c 1 -o gpurun_out/r01_v2_full -f python bench.py --steps 1 --warmup 1 --proteins-per-gpu 4000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i gpurun_out/r01_v2_full.ncu-rep --page raw --csv > gpurun_out/r01_v2_full_raw.csv 2>/dev/null
ls -la gpurun_out
