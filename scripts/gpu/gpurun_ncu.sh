ncu --set full --clock-control none --import-source on -k regex:k_score_summary_v2 -c 1 -o gpurun_out/r01_v2e_full -f python bench.py --steps 1 --warmup 1 --proteins-per-gpu 4000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i gpurun_out/r01_v2e_full.ncu-rep --page raw --csv > gpurun_out/r01_v2e_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_v2e_full.ncu-rep --page source --csv > gpurun_out/r01_v2e_full_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
