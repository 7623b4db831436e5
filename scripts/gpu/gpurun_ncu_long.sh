export PYTHONPATH=.
for n in 100000 35000; do
ncu --set full --clock-control none --import-source on -k regex:k_long_score -s 2 -c 1 -o gpurun_out/r01_long_${n} -f python scripts/gpu/long_one.py $n > /dev/null 2>&1
ncu -i gpurun_out/r01_long_${n}.ncu-rep --page raw --csv > gpurun_out/r01_long_${n}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_long_${n}.ncu-rep --page source --csv > gpurun_out/r01_long_${n}_source.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
