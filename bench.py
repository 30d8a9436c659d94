#!/usr/bin/env python
"""bench.py -- aggregate kMC steps/s of the step-loop engine on the BASELINE headline workload.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): CO oxidation on RuO2(110)
(reference: examples/render_co_oxidation_ruo2.py), 20x20 unit cells, 16384 replicas PER GPU = 16 T x 16 p_CO
grid points x 64 seeds, local_smart rule tables, default-species initial state.  A bench "step" is one
`do_kmc_steps(inner)` over the whole batch followed by the tally reduction of one sampling point (per-GPU
reduce kernel + NCCL all-reduce when N > 1).  Weak scaling: every rank owns its own 16384 replicas; there is
no inter-GPU traffic during stepping.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--inner n] [--impl ours|reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU, NCCL).
`--impl reference` times the reference semantics on the host cores (the CPU oracle port: the reference is
Fortran and this image has no Fortran compiler), process pool over all cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "aggregate kMC steps/s over replicas, RuO2 CO-ox 20x20"
UNIT = "kMC steps/s"
MODEL = "ruo2_local_smart"
SIZE = [20, 20]
REPLICAS_PER_GPU = 16384
N_T, N_P, SEEDS = 16, 16, 64


def load_workload():
    from kmos_b200 import tables, workloads
    ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", MODEL + ".json"))
    blob, info = tables.build_blob(ir)
    rates, group_of, grid = workloads.ruo2_grid(ir, n_T=N_T, n_p=N_P, seeds=SEEDS)
    assert rates.shape[0] == REPLICAS_PER_GPU
    return ir, blob, info, rates, group_of, grid


def algorithmic_bytes_per_step(P, c):
    """bytes the reference's data structures move per kMC step at reference widths.
    c = per-step averages of the oracle's event counters on this workload."""
    log2p = int(np.ceil(np.log2(P)))
    return (20 * P + 48 + 28 * P + 8 * (log2p + 3) + 8 + 32 + 12 * c["n_rs"] + 8 * c["n_chk"]
            + 36 * c["n_del"] + 8 * c["n_gs"] + 20 * c["n_add"])


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference semantics (oracle port) on all host cores, ModelRunner-style process pool
# (kmos/run/__init__.py:2330-2366): replicas dealt round-robin, each worker runs its replicas in turn.
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    blob, rates, seeds, ids, warm, n = args
    from oracle import oracle
    t_total = 0.0
    counters = np.zeros(5)
    steps = 0
    for r, seed, rid in zip(rates, seeds, ids):
        o = oracle.Oracle(blob, SIZE, seed=int(seed), replica=int(rid), rates=r)
        o.do_steps(warm)
        o.reset_counters()
        t0 = time.perf_counter()
        st = o.do_steps(n)
        t_total += time.perf_counter() - t0
        assert st == 0
        c = o.counters
        counters += [c["n_rs"], c["n_chk"], c["n_del"], c["n_gs"], c["n_add"]]
        steps += n
    return t_total, steps, counters


def cpu_pool_run(blob, rates, n_replicas, warm, n, cores, id_offset=0):
    """-> (aggregate steps/s, per-step counter averages, wall seconds of the slowest worker)"""
    import multiprocessing as mp
    pick = np.linspace(0, rates.shape[0] - 1, n_replicas).astype(int)  # spread over the (T, p) grid
    jobs = [[] for _ in range(cores)]
    for i, r in enumerate(pick):
        jobs[i % cores].append(r)
    args = [(blob, rates[j], [1000 + x for x in j], [id_offset + x for x in j], warm, n) for j in jobs if j]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(args)) as pool:
        out = pool.map(_cpu_worker, args)
    slowest = max(o[0] for o in out)
    steps = sum(o[1] for o in out)
    cnt = sum(o[2] for o in out) / steps
    counters = dict(zip(("n_rs", "n_chk", "n_del", "n_gs", "n_add"), cnt.tolist()))
    return steps / slowest, counters, slowest


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------
# the five BASELINE.json configurations (configs[2] = C is the headline; the others are reported in "configs")
# --------------------------------------------------------------------------------------------------
CONFIGS = [
    # key, label, fixture, lattice, replicas, kMC steps per timed launch
    ("A", "mini_101 fcc_100 CO adsorption/desorption, local_smart", "mini_101_local_smart", [20, 20], 16384, 20000),
    ("B", "ZGB 64x64, local_smart, y_CO sweep", "zgb_local_smart", [64, 64], 4096, 4000),
    ("C", "RuO2(110) CO oxidation 20x20, local_smart (headline)", MODEL, SIZE, REPLICAS_PER_GPU, 5000),
    ("D", "pairwise interaction 128x128, lat_int", "pairwise_lat_int", [128, 128], 2048, 4000),
    ("E", "pairwise interaction 256x256, otf", "pairwise_otf_otf", [256, 256], 3552, 40),
]
# dram bytes per launch of the dominant kernel, from the committed `ncu --set full` summaries
NCU_SUMMARIES = {"C": "ncu_summary_r2.json", "D": "ncu_latint_kernel_r1.json", "E": "ncu_otf_kernel_r1.json"}


def _counter_worker(args):
    name, size, n = args
    from kmos_b200 import tables, workloads
    from oracle import oracle
    ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
    blob, _info = tables.build_blob(ir)
    r = workloads.rates_for(name.split("_")[0], ir, 64)[32]
    o = oracle.Oracle(blob, size, seed=99, replica=0, rates=r)
    o.do_steps(n // 4)
    o.reset_counters()
    t0 = time.perf_counter()
    o.do_steps(n)
    dt = time.perf_counter() - t0
    c = o.counters
    # steps/s of the CPU port on ONE core: the `kmos benchmark` figure (README of the reference: 6.62e5 for config A)
    return {k: c[k] / float(n) for k in ("n_rs", "n_chk", "n_del", "n_gs", "n_add")}, len(ir["procs"]), n / dt


def config_counters():
    """Per-step event counters of configs A, B and D from a short oracle run each, in forked
    workers before CUDA is initialised.  -> {key: (counters, P)}"""
    import multiprocessing as mp
    jobs = [(k, (name, size, 1000000 if k == "A" else 20000)) for k, _l, name, size, _R, _n in CONFIGS if k in "ABD"]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        res = pool.map(_counter_worker, [j[1] for j in jobs])
    return {k: r for (k, _), r in zip(jobs, res)}


def _ncu_dram_bytes(key):
    try:
        with open(os.path.join(REPO, "profiles", NCU_SUMMARIES[key])) as f:
            d = json.load(f)
    except (KeyError, OSError, ValueError):
        return None, None
    if "dram_bytes_per_launch" in d:
        return d["dram_bytes_per_launch"], d.get("launch")
    conv = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    try:
        tot = sum(float(d[k]["value"].replace(",", "")) * conv[d[k]["unit"]]
                  for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        return tot, "profiles/" + NCU_SUMMARIES[key]
    except (KeyError, ValueError):
        return None, None


def measure_configs(counters, headline_entry, device, smem_peak, hbm_peak):
    """Throughput of the other four BASELINE configurations, device-resident state, CUDA events, about a
    second each, after the headline's timed region.  -> list of 5 entries (C = the headline's own numbers)."""
    from kmos_b200 import engine, otf as otf_mod, tables, workloads
    out = []
    for key, label, name, size, R, n in CONFIGS:
        if key == "C":
            out.append(headline_entry)
            continue
        ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
        m = engine.Model(ir=ir)
        rates = workloads.rates_for(name.split("_")[0], ir, R)
        lut = None
        if ir["backend"] == "otf":
            lut = np.tile(otf_mod.build_lut(ir, m.info, rates[0]), (R, 1))
        b = engine.Batch(m, R, size, device=device, rates=rates, lut=lut)
        b.do_steps(max(n // 4, 1))
        b.synchronize()
        ns0 = b.nr_of_sites.sum(axis=1).mean() if key == "E" else 0.0
        times = []
        for _ in range(2):
            b.timer_start()
            b.do_steps(n)
            times.append(b.timer_stop())
        ms = float(np.mean(times))
        info = b.kernel_info()
        ok = bool((b.status == 0).all())
        if key == "E":
            # otf: every live entry of rates_matrix is re-added every step (base_otf.f90:687-717)
            ns1 = b.nr_of_sites.sum(axis=1).mean()
            b_step = 8.0 * 0.5 * (ns0 + ns1)
        else:
            cnt, P, _cpu1 = counters[key]
            b_step = algorithmic_bytes_per_step(P, cnt)
        fast = None
        if key == "E":
            # the production selection on the same batch (kb_otf_fast.cuh: block sums, not bit-exact)
            from kmos_b200 import capi
            b.select_kernel(capi.KERNEL_OTF_FAST)
            nf = 2000
            b.do_steps(nf // 4)
            b.synchronize()
            ft = []
            for _ in range(2):
                b.timer_start()
                b.do_steps(nf)
                ft.append(b.timer_stop())
            # what a step has to read: the P row totals, the block sums of the chosen row, one block of 256
            # entries, and the event's own accesses (taken as config D's figure: the same model's lists)
            nsf = b.nr_of_sites.astype(np.float64)
            blocks = np.ceil(nsf / 256.0)
            live = nsf > 0
            sel_blocks = float(blocks[live].mean()) if live.any() else 0.0
            fast_bytes = 8.0 * (len(ir["procs"]) + sel_blocks + 256.0) + 700.0
            fast_val = R * nf / (float(np.mean(ft)) * 1e-3)
            fast = {"config": "E-fast", "workload": label + ", production selection over block sums (kb_otf_fast.cuh; "
                    "same distribution and prefix order, not bit-exact)", "model": name, "lattice": size,
                    "replicas": R, "kmc_steps_per_launch": nf, "kernel": "otf_fast",
                    "value": fast_val, "unit": UNIT, "ms_per_launch": float(np.mean(ft)),
                    "all_replicas_ok": bool((b.status == 0).all()),
                    "roofline": {"bound": "hbm", "achieved": fast_val * fast_bytes / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": fast_val * fast_bytes / 1e9 / hbm_peak if hbm_peak else None,
                                 "traffic": None, "algorithmic_bytes_per_kmc_step": fast_bytes,
                                 "note": "latency-bound: a step is a chain of dependent DRAM round trips over 17 GB "
                                         "of state (profiles/ncu_otf_fast_kernel_lines_r2b.txt), not a stream"}}
        b.close()
        m.close()
        launch_bytes = b_step * R * n
        achieved = launch_bytes / (ms * 1e-3) / 1e9
        in_smem = info["kernel_name"] in ("generated", "smem")
        peak = smem_peak if in_smem else hbm_peak
        traffic, traffic_src = _ncu_dram_bytes(key)
        out.append({
            "config": key, "workload": label, "model": name, "lattice": size, "replicas": R,
            "kmc_steps_per_launch": n, "kernel": info["kernel_name"], "replicas_per_cta": info["replicas_per_cta"],
            "value": R * n / (ms * 1e-3), "unit": UNIT, "ms_per_launch": ms, "all_replicas_ok": ok,
            "roofline": {"bound": "smem" if in_smem else "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_kmc_step": b_step}})
        if key in counters:
            out[-1]["cpu_port_single_core"] = {"value": counters[key][2], "unit": UNIT,
                                               "note": "one replica on one host core, the `kmos benchmark` shape"}
        if fast is not None:
            out.append(fast)
    return out


# --------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    ir, blob, info, rates, group_of, grid = load_workload()
    cores = host_cores()
    n_rep = 4 * cores
    warm, n = 20000, args.cpu_steps
    # untimed warm-up steps, then K timed steps; each step = the bounded sample below
    vals = []
    for i in range(args.warmup + args.steps):
        v, _c, wall = cpu_pool_run(blob, rates, n_rep, warm, n, cores)
        if i >= args.warmup:
            vals.append((v, wall))
    value = float(np.mean([v for v, _ in vals]))
    sample = "%d replicas spread over the T x p_CO grid x %d steps each (after %d warm-up steps), %d workers" % (
        n_rep, n, warm, cores)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([w for _, w in vals]) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": workload_config(grid, args),
        "note": "reference semantics on host cores (CPU port of the generated Fortran, gcc -O3; neither this "
                "image nor the GPU box has a Fortran compiler: profiles/fortran_probe_r2.txt)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(grid, args):
    """The same dictionary in both arms: what is computed, not how."""
    return {"workload": "RuO2(110) CO oxidation 20x20 (examples/render_co_oxidation_ruo2.py), local_smart, "
                        "%d replicas/GPU = %d T x %d p_CO x %d seeds" % (REPLICAS_PER_GPU, N_T, N_P, SEEDS),
            "replicas_per_gpu": REPLICAS_PER_GPU, "lattice": SIZE, "processes": 36,
            "kmc_steps_per_replica_per_step": args.inner,
            "T_K": [grid["T"][0], grid["T"][-1]], "p_CO_bar": [grid["p_COgas"][0], grid["p_COgas"][-1]],
            "p_O2_bar": grid["p_O2gas"], "rng": "Philox4x32-10 per replica",
            "parallelism": "replica-sharded, %d GPU(s), tally all-reduce per step (int64 counts + f64 sums)" % args.gpus,
            # timing rule: inputs larger than L2 -- every bench step streams each replica's whole state
            # (avail_sites incl. the L2-resident lists, lattice, per-process arrays: > 28 KB per replica)
            "l2": "no flush: the state a step touches (> %d MB/GPU) exceeds the 126 MB L2"
                  % (REPLICAS_PER_GPU * 28 * 1024 // 2**20)}


def run_ours(args, rank, world, local_rank):
    ir, blob, info, rates, group_of, grid = load_workload()
    P = len(ir["procs"])

    # ---- CPU legs first (before CUDA is initialised in this process: the pools fork) --------------------
    cpu = None
    counters = None
    cfg_counters = None
    if rank == 0:
        cores = host_cores()
        n_rep, warm, n = 4 * cores, 20000, args.cpu_steps
        v, counters, wall = cpu_pool_run(blob, rates, n_rep, warm, n, cores)
        if world == 1:
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d replicas spread over the grid x %d steps (after %d warm-up), %.1f s wall"
                             % (n_rep, n, warm, wall)}
            if not args.no_configs:
                cfg_counters = config_counters()

    import torch
    from kmos_b200 import engine, parallel

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()
    model = engine.Model(ir=ir, blob=blob, info=info)
    n_groups = N_T * N_P
    nocc = model.n_species * model.spuck
    count_cols = parallel.count_columns(P, nocc)
    smem_peak, _mhz = engine.measure_smem_bandwidth(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_batch(first, count):
        """Replicas [first, first+count) of the global id space, their rows of the rate matrix."""
        gid = np.arange(first, first + count)
        seeds = gid.astype(np.uint64) * np.uint64(2654435761) + np.uint64(17)
        rows = np.ascontiguousarray(rates[gid % REPLICAS_PER_GPU])
        b = engine.Batch(model, count, SIZE, device=local_rank, seeds=seeds, replica_ids=gid.astype(np.uint32), rates=rows)
        b.set_stream(stream.cuda_stream)
        return b, np.ascontiguousarray(group_of[gid % REPLICAS_PER_GPU]), rows

    def timed_region(batch, groups, tally, steps, warmup, kernel_events=None):
        def one_step(kev=None):
            with torch.cuda.stream(stream):
                if kev is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                batch.do_steps(args.inner)
                if kev is not None:
                    e1.record(stream)
                    kev.append((e0, e1))
                batch.reduce_tallies(groups, n_groups, dev_ptr=tally.data_ptr(), want_host=False)
                parallel.all_reduce_tallies(tally, count_cols=count_cols)
        for _ in range(warmup):
            one_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record(stream)
        for _ in range(steps):
            one_step(kernel_events)
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- weak scaling: every rank owns REPLICAS_PER_GPU replicas; device-resident throughput ("value") ---
    batch, groups, rate_rows = make_batch(rank * REPLICAS_PER_GPU, REPLICAS_PER_GPU)
    kinfo = batch.kernel_info()
    assert kinfo["kernel_name"] == "generated", kinfo
    words = batch.tally_words()
    tally = torch.zeros((n_groups, words), dtype=torch.float64, device="cuda")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kev = []
    ms = timed_region(batch, groups, tally, args.steps, args.warmup, kev)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms = float(t.item())
    status_ok = bool(np.all(batch.status == 0))
    steps_done = int(batch.kmc_step.min())
    # the all-reduced tally of the last step must account for every replica and step of every rank
    tl = batch.split_tally(tally.cpu().numpy())
    checks = {"all_replicas_ok": status_ok, "kmc_steps_per_replica_total": steps_done,
              "tally_n_replicas": float(tl["n_replicas"].sum()), "tally_kmc_steps": float(tl["kmc_steps"].sum()),
              "tally_events": float(tl["procstat"].sum())}
    assert checks["tally_n_replicas"] == world * REPLICAS_PER_GPU, checks
    if status_ok:
        assert checks["tally_kmc_steps"] == checks["tally_events"] == float(world) * REPLICAS_PER_GPU * steps_done, checks

    # ---- end to end through the public API with host buffers ("e2e") -------------------------------------
    pinned = torch.from_numpy(np.ascontiguousarray(rate_rows)).pin_memory()
    rates_pinned = pinned.numpy()
    host_tally = np.zeros((n_groups, words))

    def one_step_e2e():
        batch.set_rates(rates_pinned)                       # H2D of this step's inputs (pinned host memory)
        batch.do_steps(args.inner)
        t = batch.reduce_tallies(groups, n_groups, dev_ptr=tally.data_ptr(), want_host=(world == 1))
        if world > 1:
            with torch.cuda.stream(stream):
                parallel.all_reduce_tallies(tally, count_cols=count_cols)
                t = tally.cpu().numpy()                      # D2H of the step's result
            stream.synchronize()
        host_tally[:] = t

    for _ in range(max(1, args.warmup // 2)):
        one_step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        one_step_e2e()
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    batch.close()

    # ---- strong scaling: the same 16384 replicas split over the ranks --------------------------------------
    total_steps = float(REPLICAS_PER_GPU) * args.inner * args.steps * world
    strong = None
    if world > 1:
        lo, hi = parallel.shard_bounds(REPLICAS_PER_GPU, rank, world)
        sb, sgroups, _rows = make_batch(lo, hi - lo)
        sms = timed_region(sb, sgroups, tally, args.steps, args.warmup)
        stl = sb.split_tally(tally.cpu().numpy())
        assert float(stl["n_replicas"].sum()) == REPLICAS_PER_GPU
        strong = {"value": float(REPLICAS_PER_GPU) * args.inner * args.steps / (sms * 1e-3), "unit": UNIT,
                  "ms_per_step": sms / args.steps, "replicas_total": REPLICAS_PER_GPU,
                  "replicas_per_gpu": hi - lo, "kernel": sb.kernel_info()["kernel_name"],
                  "replicas_resident_per_gpu": sb.kernel_info()["replicas_per_cta"] * sb.kernel_info()["ctas_per_sm"]
                  * sb.kernel_info()["sm_count"]}
        sb.close()
    else:
        strong = {"value": total_steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
                  "replicas_total": REPLICAS_PER_GPU, "replicas_per_gpu": REPLICAS_PER_GPU,
                  "note": "N = 1: identical to the weak-scaling run"}

    if rank == 0:
        value = total_steps / (ms * 1e-3)
        b_step = algorithmic_bytes_per_step(P, counters)
        launch_bytes = b_step * REPLICAS_PER_GPU * args.inner
        achieved = launch_bytes / (kernel_ms * 1e-3) / 1e9
        peaks = {}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = _ncu_dram_bytes("C")
        # The kernel keeps its hot state (class planes, lattice, list windows, nr_of_sites) in shared memory,
        # so the roofline that bounds it is the shared-memory one; peak = LDS.128 streaming
        # microbenchmark measured live on this GPU.  The HBM view is kept alongside under "hbm".
        roofline = {
            "bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s",
            "frac": achieved / smem_peak if smem_peak else None,
            "peak_source": "live LDS.128 streaming microbenchmark on this GPU (kmos_b200_measure_smem_bandwidth)",
            "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "kb_gen_kernel<ruo2>", "kernel_ms_per_launch": kernel_ms,
            "kernel_share_of_step": kernel_ms * args.steps / ms,
            "algorithmic_bytes_per_launch": launch_bytes, "algorithmic_bytes_per_kmc_step": b_step,
            "event_counters_per_step": counters,
            "hbm": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured, burst copy)" if peaks else "fallback 6650 GB/s"},
            "note": "algorithmic bytes are counted at the reference's data widths (SURVEY 8d); the kernel is "
                    "latency-bound: one replica is a serial dependency chain (DESIGN.md 4.1)",
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(grid, args),
            "kernel": kinfo, "checks": checks,
            "clocks": clocks,
            "e2e": {"value": total_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(rate_rows.nbytes), "d2h_bytes_per_step": int(host_tally.nbytes),
                    "timing": "host wall clock, synchronized both sides, max over ranks"},
            "gpu_launches": 3 * args.steps,
            "roofline": roofline,
            "strong": strong,
        }
        if args.scaling == "strong":  # report the strong-scaling run on top, keep the weak one beside it
            line["weak"] = {"value": line["value"], "unit": UNIT, "ms_per_step": line["ms_per_step"],
                            "replicas_per_gpu": REPLICAS_PER_GPU}
            line["value"], line["ms_per_step"], line["scaling"] = strong["value"], strong["ms_per_step"], "strong"
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if cfg_counters is not None:
            head = {"config": "C", "workload": CONFIGS[2][1], "model": MODEL, "lattice": SIZE,
                    "replicas": REPLICAS_PER_GPU, "kmc_steps_per_launch": args.inner, "kernel": kinfo["kernel_name"],
                    "replicas_per_cta": kinfo["replicas_per_cta"], "value": value, "unit": UNIT,
                    "ms_per_launch": kernel_ms, "all_replicas_ok": status_ok,
                    "roofline": {k: roofline[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic",
                                                          "traffic_source", "algorithmic_bytes_per_kmc_step")}}
            line["configs"] = measure_configs(cfg_counters, head, local_rank, smem_peak, hbm_peak)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_config(args, key, rank, world, local_rank):
    """One of the other BASELINE configurations (A, B, D, E) as its own weak-scaling job: the configuration's
    replica count PER GPU, replicas sharded over the ranks by global id, one NCCL tally all-reduce per step
    (`--config E --gpus 8` is BASELINE.json configs[4]).  Same JSON contract, without the CPU arm."""
    import torch
    from kmos_b200 import engine, otf as otf_mod, parallel, tables, workloads
    _k, label, name, size, R, inner = [c for c in CONFIGS if c[0] == key][0]
    inner = args.inner if args.inner != 5000 else inner
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()
    ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
    model = engine.Model(ir=ir)
    P = model.n_proc
    rates = workloads.rates_for(name.split("_")[0], ir, R)
    lut = None
    if ir["backend"] == "otf":
        lut = np.tile(otf_mod.build_lut(ir, model.info, rates[0]), (R, 1))
    gid = np.arange(rank * R, (rank + 1) * R)
    seeds = gid.astype(np.uint64) * np.uint64(2654435761) + np.uint64(17)
    batch = engine.Batch(model, R, size, device=local_rank, seeds=seeds, replica_ids=gid.astype(np.uint32),
                         rates=rates, lut=lut)
    batch.set_stream(stream.cuda_stream)
    n_groups = 8
    groups = (np.arange(R) * n_groups // R).astype(np.int32)
    words = batch.tally_words()
    tally = torch.zeros((n_groups, words), dtype=torch.float64, device="cuda")
    count_cols = parallel.count_columns(P, model.n_species * model.spuck)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        with torch.cuda.stream(stream):
            batch.do_steps(inner)
            batch.reduce_tallies(groups, n_groups, dev_ptr=tally.data_ptr(), want_host=False)
            parallel.all_reduce_tallies(tally, count_cols=count_cols)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    for _ in range(args.steps):
        one_step()
    t1.record(stream)
    barrier()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ok = bool(np.all(batch.status == 0))
    steps_done = int(batch.kmc_step.min())
    tl = batch.split_tally(tally.cpu().numpy())
    checks = {"all_replicas_ok": ok, "kmc_steps_per_replica_total": steps_done,
              "tally_n_replicas": float(tl["n_replicas"].sum()), "tally_kmc_steps": float(tl["kmc_steps"].sum())}
    assert checks["tally_n_replicas"] == world * R, checks
    if ok:
        assert checks["tally_kmc_steps"] == float(world) * R * steps_done, checks
    kinfo = batch.kernel_info()
    ns = float(batch.nr_of_sites.sum(axis=1).mean())
    batch.close()
    if rank == 0:
        total = float(R) * inner * args.steps * world
        line = {"metric": "aggregate kMC steps/s over replicas, %s" % label, "value": total / (ms * 1e-3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": label, "baseline_config": key, "model": name, "lattice": size,
                           "replicas_per_gpu": R, "kmc_steps_per_replica_per_step": inner,
                           "parallelism": "replica-sharded, %d GPU(s), tally all-reduce per step (int64 counts + f64 sums)" % world,
                           "l2": "no flush: the state a step touches exceeds the 126 MB L2" if key in "DE" else
                                 "state per GPU: %d replicas" % R},
                "kernel": kinfo, "checks": checks, "clocks": clocks, "gpu_launches": 3 * args.steps}
        if key == "E":
            b_step = 8.0 * ns
            peaks = {}
            try:
                with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
                    peaks = json.load(f)
            except (OSError, ValueError):
                pass
            hbm = peaks.get("hbm_gbs", 6650.0)
            ach = b_step * R * inner * args.steps / (ms * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                                "traffic": _ncu_dram_bytes("E")[0], "algorithmic_bytes_per_kmc_step": b_step,
                                "note": "per GPU; every live rates_matrix entry is re-added every step"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--inner", type=int, default=5000, help="kMC steps per replica per bench step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=800000, help="kMC steps per replica of the CPU sample")
    ap.add_argument("--no-configs", action="store_true", help="skip the other four BASELINE configurations")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which run the line's value / ms_per_step / scaling report: weak (16384 replicas per GPU, "
                         "default) or strong (16384 replicas in total); the other one is always there as a sub-object")
    ap.add_argument("--config", default="C", choices=["A", "B", "C", "D", "E"],
                    help="BASELINE configuration to run as the job (default C, the headline)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config != "C":
        run_config(args, args.config, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


def _emit(line):
    """The one JSON line of the contract goes to the process' original stdout."""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    # Libraries (NCCL with NCCL_DEBUG=VERSION/INFO in the environment, torch warnings) print to fd 1; the
    # contract is exactly one JSON line on stdout, so everything else is sent to stderr.
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    main()
