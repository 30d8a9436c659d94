#!/bin/bash
# A/B of generated-kernel variants on the headline workload (how the round-2 experiments in DESIGN.md 4.1 were run):
#   bash tools/gpu_ab.sh <lanes per replica> "<-D switches of variant 1>" "<variant 2>" ...
# every variant is a separately generated + compiled module (KMOS_B200_GEN_DEFS is part of the cache key)
mkdir -p gpurun_out
L=${1:-16}
shift
i=0
for D in "$@"; do
  i=$((i+1))
  KMOS_B200_GEN_LPR=$L KMOS_B200_GEN_DEFS="$D" timeout 600 python tools/gen_probe.py ruo2gen > gpurun_out/ab_$i.log 2>&1
  echo "== [$D]"; grep -o '"kernel": "[a-z]*", "ms": [0-9.]*, "steps_per_s": [0-9.]*' gpurun_out/ab_$i.log; grep -i "error\|fail" gpurun_out/ab_$i.log | head -3
done
