#!/bin/bash
# final evidence of the round: bench line, launch list, --set full capture of the headline kernel at the bench shape
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; tail -2 gpurun_out/bench_n1_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1_final.json 2>/dev/null
bash tools/gpu_prof_r2.sh
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1_final.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
for c in d["configs"]: print(c["config"], c["kernel"], "%.3e" % c["value"], c.get("roofline",{}).get("frac"))
print(json.load(open("gpurun_out/bench_ref_n1_final.json"))["value"])
PY
