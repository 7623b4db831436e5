for m in 2 3 4; do
PLAAC_T3_MINB=$m PLAAC_TRACKS=3 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_res_tracks3 -c 2 --csv --log-file gpurun_out/t3_$m.csv python scripts/gpu/res_once.py > /dev/null 2>&1
echo "minb $m: $(grep k_res_tracks3 gpurun_out/t3_$m.csv | tail -1 | awk -F'","' '{print $NF}')"
done
