#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
P=29811
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config E --steps 5 --warmup 3 > gpurun_out/bench_otf_n${N}_r2.json 2> gpurun_out/bench_otf_n${N}_r2.err
tail -2 gpurun_out/bench_otf_n${N}_r2.err; cut -c1-400 gpurun_out/bench_otf_n${N}_r2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err
tail -2 gpurun_out/bench_n${N}_r2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}_r2.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus","checks","strong")}); print(d["e2e"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n${N}_r2.json 2>/dev/null
cut -c1-300 gpurun_out/bench_ref_n${N}_r2.json
