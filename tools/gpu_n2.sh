#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -5
P=29711
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err
tail -2 gpurun_out/bench_n2_r2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2_r2.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus","checks","strong")}); print(d["e2e"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --config E --steps 5 --warmup 3 > gpurun_out/bench_otf_n2_r2.json 2> gpurun_out/bench_otf_n2_r2.err
tail -2 gpurun_out/bench_otf_n2_r2.err; cut -c1-600 gpurun_out/bench_otf_n2_r2.json
