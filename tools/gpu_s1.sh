#!/bin/bash
# round-2 session: toolchain probe, GPU suite, generated vs interpreter, source-level ncu captures
mkdir -p gpurun_out
{ which gfortran gfortran-13 gfortran-12 flang-new flang nvfortran f95 f77; ls /usr/lib/gcc/*/*/f951 /usr/libexec/gcc/*/*/f951; nproc; } > gpurun_out/fortran_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/gputest.log
timeout 600 python tools/gen_probe.py ruo2 > gpurun_out/gen_ruo2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:kb_smem_kernel -s 1 -c 1 -f -o gpurun_out/smem_r2a python tools/ncu_probe.py ruo2_local_smart 16384 1000 > gpurun_out/ncu_smem.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:kb_gen_kernel -s 1 -c 1 -f -o gpurun_out/gen_r2a python tools/ncu_probe.py ruo2_local_smart 16384 1000 20x20 generated > gpurun_out/ncu_gen.log 2>&1
tail -3 gpurun_out/gputest.log; cat gpurun_out/gen_ruo2.log | cut -c1-300; cat gpurun_out/fortran_probe.txt
