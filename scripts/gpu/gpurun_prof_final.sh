# round-end captures of the final build: the dominant kernel (full set, the bench shard) and the per-residue launch list
ncu --set full --clock-control none --import-source on -k regex:k_score_summary_v2 -c 1 -o gpurun_out/r01_v2h_full_12m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > /dev/null 2>&1
ncu -i gpurun_out/r01_v2h_full_12m.ncu-rep --page raw --csv > gpurun_out/r01_v2h_full_12m_raw.csv 2>/dev/null
bash scripts/gpu/gpurun_res_launches.sh > gpurun_out/r01_res_launches_summary.txt 2>&1
tail -20 gpurun_out/r01_res_launches_summary.txt
ls -la gpurun_out | tail -6
