ncu --set full --clock-control none -k regex:k_pack -c 1 -o gpurun_out/r01_pack_full -f python bench.py --steps 1 --warmup 1 --proteins-per-gpu 4000000 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > /dev/null 2>&1
ncu -i gpurun_out/r01_pack_full.ncu-rep --page raw --csv > gpurun_out/r01_pack_full_raw.csv 2>/dev/null
python profiles/summarise_ncu.py gpurun_out/r01_pack_full_raw.csv 1406000000
