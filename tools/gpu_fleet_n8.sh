#!/bin/bash
# 8-GPU box: one process driving all eight GPUs through the fleet, then the N=8 bench line (own arm).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 120 python tools/fleet_probe.py 8 5000 5 2> gpurun_out/fleet_probe_n8.err | tee gpurun_out/fleet_probe_n8.json
tail -3 gpurun_out/fleet_probe_n8.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 10 --warmup 3 --no-configs > gpurun_out/bench_n8_final.json 2> gpurun_out/bench_n8_final.err
tail -2 gpurun_out/bench_n8_final.err; cut -c1-300 gpurun_out/bench_n8_final.json
