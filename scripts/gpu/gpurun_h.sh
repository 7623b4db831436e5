set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -c 400 gpurun_out/bench_h.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_h.json 2>> gpurun_out/bench_h.err
python -c "
import json; d=json.load(open('gpurun_out/bench_h.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']); print(d['e2e']); print(d['cpu_baseline']); print(d['extras']['long_sequences'])"
