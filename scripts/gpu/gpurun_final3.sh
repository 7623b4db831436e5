set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_final.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2>> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-per-residue --no-extras > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['gpu_launches']); print(d['e2e']); print(d['cpu_baseline']['value'], d['cpu_baseline']['gpu_parity_on_sample']['ok']); print(d['extras']['long_sequences']); print(d['per_residue_mode']['yeast_sized_6k']['ms'], d['per_residue_mode']['batch_200k']['ms'])"
