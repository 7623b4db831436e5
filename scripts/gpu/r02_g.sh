PLAAC_TRACKS=3 REPS=1 ncu --set full --clock-control none --import-source on -k regex:k_res_tracks3 -c 1 -o gpurun_out/r02_tracks3 -f python scripts/gpu/res_once.py > /dev/null 2>&1
ncu -i gpurun_out/r02_tracks3.ncu-rep --page raw --csv > gpurun_out/r02_tracks3_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_tracks3.ncu-rep --page source --csv > gpurun_out/r02_tracks3_source.csv 2>/dev/null
ls -la gpurun_out/r02_tracks3*
