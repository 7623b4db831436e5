B="python bench.py --steps 5 --warmup 3 --proteins-per-gpu 4000000 --no-cpu-baseline --no-e2e --no-extras --no-per-residue"
P="import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g kernel_ms %.2f frac %.3f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
cp plaac_b200/csrc/summary_kernel_v2.cuh /tmp/orig.cuh
for T in 768 640; do
  cp /tmp/orig.cuh plaac_b200/csrc/summary_kernel_v2.cuh
  sed -i "s/^constexpr int kV2MaxThreads = [0-9]*;/constexpr int kV2MaxThreads = $T;/" plaac_b200/csrc/summary_kernel_v2.cuh
  PYTHONPATH=. python scripts/gpu/gpurun_tb_patch.py
  python plaac_b200/build.py -f -v 2>&1 | grep -A3 "k_score_summary_v2" | grep "Used"
  echo "threads=$T tb-patch"; $B | python -c "$P"
done
cp /tmp/orig.cuh plaac_b200/csrc/summary_kernel_v2.cuh
sed -i "s/^constexpr int kV2MaxThreads = [0-9]*;/constexpr int kV2MaxThreads = 640;/" plaac_b200/csrc/summary_kernel_v2.cuh
python plaac_b200/build.py -f
echo "threads=640 plain, 12.5M"; python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-per-residue | python -c "$P"
