#!/bin/bash
# 2-GPU box: the fleet (one process, two devices) against the oracle and the single batch, the NCCL multi-rank
# tests, the fleet probe, the N=2 bench line.
mkdir -p gpurun_out
nvidia-smi -L | head -4
(timeout 300 python -m pytest tests/test_gpu_fleet.py tests/test_gpu_multi.py -q 2>&1 | tail -6) | tee gpurun_out/fleet_n2_test.log
timeout 200 python tools/fleet_probe.py 2 5000 5 2> gpurun_out/fleet_probe_n2.err | tee gpurun_out/fleet_probe_n2.json
tail -3 gpurun_out/fleet_probe_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 3 --no-configs > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
tail -2 gpurun_out/bench_n2_final.err; cut -c1-400 gpurun_out/bench_n2_final.json
