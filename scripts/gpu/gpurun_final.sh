set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 400 gpurun_out/bench_f.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_f.json 2>> gpurun_out/bench_f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01_v2g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_summary_v2 -c 1 -o gpurun_out/r01_v2g_full_12m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > /dev/null 2>&1
ncu -i gpurun_out/r01_v2g_full_12m.ncu-rep --page raw --csv > gpurun_out/r01_v2g_full_12m_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
