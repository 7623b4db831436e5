# round-2 captures of the final build: launch list of the bench, the dominant kernel (full set, the bench shard), the
# long per-residue kernel (full set, one 100 k-residue protein) and the per-residue launch list
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > gpurun_out/r02_final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_summary_v3 -c 1 -o gpurun_out/r02_v3_full_12m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > /dev/null 2>&1
ncu -i gpurun_out/r02_v3_full_12m.ncu-rep --page raw --csv > gpurun_out/r02_v3_full_12m_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_long_post -c 1 -o gpurun_out/r02_longpost_100000 -f python scripts/gpu/r02_longres_one.py 100000 > /dev/null 2>&1
ncu -i gpurun_out/r02_longpost_100000.ncu-rep --page raw --csv > gpurun_out/r02_longpost_100000_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_res_launches.csv python scripts/gpu/res_once.py > /dev/null 2>&1
rm -f gpurun_out/r02_v3_full_12m.ncu-rep.tmp
ls -la gpurun_out | tail -8
