#!/bin/bash
mkdir -p gpurun_out
for L in 16 8 32; do
  KMOS_B200_GEN_LPR=$L timeout 600 python tools/gen_probe.py parity ruo2 > gpurun_out/gen_lpr$L.log 2>&1
  echo "== LPR $L"; cut -c1-260 gpurun_out/gen_lpr$L.log
done
