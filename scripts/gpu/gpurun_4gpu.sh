timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu_k.json 2> gpurun_out/bench_4gpu_k.err
echo "stdout lines: $(wc -l < gpurun_out/bench_4gpu_k.json)"; python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu_k.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['clocks']); print(d['e2e'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_ref_4gpu_k.json 2>> gpurun_out/bench_4gpu_k.err
echo "ref stdout lines: $(wc -l < gpurun_out/bench_ref_4gpu_k.json)"; head -c 300 gpurun_out/bench_ref_4gpu_k.json
