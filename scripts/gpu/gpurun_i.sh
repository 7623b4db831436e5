python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-per-residue > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -c 300 gpurun_out/bench_i.err
python -c "
import json; d=json.load(open('gpurun_out/bench_i.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']); print(d['e2e']['ms_per_step']); print(d['extras']['ranking'])"
