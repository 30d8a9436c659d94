#!/bin/bash
mkdir -p gpurun_out
L=${1:-16}
KMOS_B200_GEN_LPR=$L timeout 900 ncu --set full --import-source on --clock-control none -k regex:kb_gen_kernel -s 1 -c 1 -f -o gpurun_out/gen_lpr${L}_r2 python tools/ncu_probe.py ruo2_local_smart 16384 1000 20x20 generated > gpurun_out/ncu_gen_lpr$L.log 2>&1
tail -3 gpurun_out/ncu_gen_lpr$L.log
