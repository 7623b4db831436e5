# one 8-GPU call on the final round-2 build: multi-GPU tests on distinct devices, bench at N=8 incl. the single-process
# multi_ctx side measurement (plaac_score_multi_packed: hit rows merged on one GPU over peer copies)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu_b.json 2> gpurun_out/r02_bench_8gpu_b.err
tail -2 gpurun_out/r02_bench_8gpu_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_8gpu_b.json'))
print('N=8 value %.4g  e2e %.4g (%.1f ms)  variants: %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], {k: '%.3g' % v['value'] for k, v in d['e2e']['variants'].items()}))
print('multi_ctx', d.get('multi_ctx_e2e'))
PY
