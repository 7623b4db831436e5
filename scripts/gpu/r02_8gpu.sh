# one 8-GPU call: raw link ceiling with 1/2/4/8 ranks at once, the multi-GPU tests on distinct devices, bench at N=8 (incl. the
# single-process multi_ctx side measurement) and at N=2
mkdir -p gpurun_out
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
      scripts/gpu/link_probe.py --out gpurun_out/r02_link_probe_$n.json 2>/dev/null | tail -1
done
python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -2 gpurun_out/r02_bench_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_8gpu.json'))
print('N=8 value %.4g  e2e %.4g (%.1f ms)  variants: %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], {k: '%.3g' % v['value'] for k, v in d['e2e']['variants'].items()}))
print('multi_ctx', d.get('multi_ctx_e2e'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print('N=2 value %.4g  e2e %.4g (%.1f ms)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('multi_ctx', d.get('multi_ctx_e2e'))
PY
