for T in 768 704 640 576; do
  sed -i "s/^constexpr int kV2MaxThreads = [0-9]*;/constexpr int kV2MaxThreads = $T;/" plaac_b200/csrc/summary_kernel_v2.cuh
  python plaac_b200/build.py -f -v 2>&1 | grep -A3 "k_score_summary_v2" | grep "Used" 
  echo "threads=$T"; python bench.py --steps 5 --warmup 3 --proteins-per-gpu 4000000 --no-cpu-baseline --no-e2e --no-extras --no-per-residue | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.2f kernel_ms %.2f frac %.3f share %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['roofline']['kernel_share_of_step']))"
done
