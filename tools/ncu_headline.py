#!/usr/bin/env python
"""profiles/ncu_gen_kernel_r2.json (tools/ncu_summarize.py full ...) -> profiles/ncu_summary_r2.json, the
handful of numbers bench.py and DESIGN.md quote (dram bytes per launch = roofline.traffic).
    python tools/ncu_headline.py <full json> <summary json> [replicas] [steps per launch]"""
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
R = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
N = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
d = json.load(open(src))


def f(k):
    return float(d[k]["value"].replace(",", ""))


conv = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
rd = f("dram__bytes_read.sum") * conv[d["dram__bytes_read.sum"]["unit"]]
wr = f("dram__bytes_write.sum") * conv[d["dram__bytes_write.sum"]["unit"]]
out = {
    "kernel": d["Kernel Name"]["value"],
    "command": "ncu --set full --clock-control none --import-source on -k regex:kb_gen_kernel -s 3 -c 1 "
               "python bench.py --steps 1 --warmup 3 --cpu-steps 20000 --no-configs",
    "launch": ("%d replicas x %d kMC steps; " % (R, N)) + "grid %s x %s threads; %s %s dynamic smem/CTA" % (
        d["launch__grid_size"]["value"], d["launch__block_size"]["value"],
        d["launch__shared_mem_per_block_dynamic"]["value"], d["launch__shared_mem_per_block_dynamic"]["unit"]),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "duration_ms": f("gpu__time_duration.sum"),
    "ipc_per_sm_active": f("sm__inst_executed.avg.per_cycle_active"),
    "warps_active_per_sm": f("sm__warps_active.avg.per_cycle_active"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": int(f("launch__registers_per_thread")),
    "instructions_per_launch": f("smsp__inst_executed.sum"),
    "instructions_per_kmc_step": f("smsp__inst_executed.sum") / (R * N),
    "smem_wavefronts_pct_of_peak": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "warp_latency_per_inst_issued_cycles": f("smsp__average_warp_latency_per_inst_issued.ratio"),
    "stall_wait": f("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    "stall_short_scoreboard": f("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    "stall_long_scoreboard": f("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    "stall_not_selected": f("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    "stall_branch_resolving": f("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    "stall_math_pipe_throttle": f("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    "note": "absolute duration under ncu is serialised/cold; bench.py reports the live CUDA-event time",
}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
