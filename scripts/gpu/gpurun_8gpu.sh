free -g | head -2; nproc; nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 600 gpurun_out/bench_8gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['clocks']); print(d['e2e'])"
