python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -2 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print('N=2 value %.4g  e2e %.4g (%.1f ms)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('multi_ctx', d.get('multi_ctx_e2e'))
PY
