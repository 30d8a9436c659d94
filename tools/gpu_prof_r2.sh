#!/bin/bash
# round-2 evidence: launch list of a short bench run, --set full capture of the headline kernel at the bench shape
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --cpu-steps 20000 --no-configs > gpurun_out/launches_r2.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kb_gen_kernel -s 3 -c 1 -f -o gpurun_out/gen_headline_r2 python bench.py --steps 1 --warmup 3 --cpu-steps 20000 --no-configs > gpurun_out/ncu_headline_r2.out 2>&1
tail -2 gpurun_out/ncu_headline_r2.out | cut -c1-300
