#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/soak.py > gpurun_out/soak_r2.log 2>&1; tail -2 gpurun_out/soak_r2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py > gpurun_out/memcheck_r2.log 2>&1; echo "memcheck rc $?"; grep -c "ERROR SUMMARY" gpurun_out/memcheck_r2.log; grep "ERROR SUMMARY" gpurun_out/memcheck_r2.log | tail -2
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_probe.py > gpurun_out/racecheck_r2.log 2>&1; echo "racecheck rc $?"; grep "RACECHECK SUMMARY\|hazard" gpurun_out/racecheck_r2.log | sort | uniq -c | sort -rn | head -8
