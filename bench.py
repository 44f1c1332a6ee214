#!/usr/bin/env python
"""bench.py -- DP GCUPS of the B200 alignment-DP engine on BASELINE.json config 2
("dynprog microbench: 1M synthetic DP boxes, 50-2000 bp per side, single/cdna/genome/end5/end3").

  python bench.py --gpus 1 --steps K --warmup W             our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus 1 --steps K ...   the reference's own CPU DP on the host cores

A step = one pass of the hot path over the rank's batch of synthetic boxes (weak scaling: every rank
gets its own `--boxes` boxes).  `value` = algorithmic in-band cells of all ranks / device time of a
step (max over ranks), inputs resident in HBM.  `e2e` = the same through the entry-point API
(GmapDP_batch_run): H2D of boxes + sequences, kernels, D2H of results + edit scripts AND the replay of
every edit script into the pair list the reference's entry point returns, all inside the timed region;
`e2e_device` stops at the device ABI (gmapdp_run_batch, no pair lists).  Cells are counted as
SURVEY.md section 8(d) defines them; calls that the reference resolves without a fill
(single_gap_simple, genome_gap_simple, QUERYEND_NOGAPS, require_pos_score_p) count zero cells.
Other keys of the line: `strata.production` (the same measurement on 15-150 bp boxes, > 95 % of GMAP's own
calls), `chain` (second kernel), `program` (whole program: cDNAs/s of gmap.sm100 beside the stock gmap.avx2
on a scaled BASELINE config 3), `cpu_baseline` (+ its no-clflush fairness variant).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "dp_gcups"
UNIT = "GCUPS"
WORKLOAD = "dynprog microbench: synthetic DP boxes, 50-2000 bp per side, modes single/genome/cdna/end5/end3 in fifths (BASELINE.json configs[1])"
BYTES_PER_CELL_8, BYTES_PER_CELL_16 = 2, 3          # score + packed directions
OPS_PER_CELL_TRI, OPS_PER_CELL_FULL = 12, 18        # algorithmic integer ops


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def int_peak():
    """integer issue rate measured on this pool's B200 by scripts/int_peak.cu (profiles/int_peak.json), Tera lane-ops/s;
    fallback = 148 SMs x 128 lanes/clk x 1.965 GHz"""
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "int_peak.json")))
        return float(pk["iadd3"]), "measured: profiles/int_peak.json (scripts/int_peak.cu, dependent-free IADD3 chains = 128 lanes/clk/SM; " \
            "fused add+max forms VIADDMNMX / VIMNMX3 issue at %.1f)" % pk["viaddmnmx_s32"]
    except Exception:
        return 148 * 128 * 1.965e9 / 1e12, "nominal: 148 SMs x 128 int32 lanes/clk x 1.965 GHz"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own DP on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, seed, start, stride, budget_s, small, resident = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import benchgen
    if resident:
        benchgen.resident_tables_only(seed)     # the boxes' probability arrays = the MaxEnt values of the resident form
    from harness import Oracle, Ref
    counter = Oracle()                      # cells are counted by the oracle's own entry-point control flow, not by the product
    runner = Oracle() if kind == "port" else Ref(noflush=(kind == "reference_noflush"))
    is_ref = kind != "port"
    t0 = time.perf_counter()
    i, nboxes, cells, busy = start, 0, 0, 0.0
    while True:
        box = benchgen.make(seed, i, small)
        cells += counter.count_cells(box)
        if is_ref:
            box = benchgen.attach_ref_world(runner, box)
        t1 = time.perf_counter()
        runner.run(box)
        busy += time.perf_counter() - t1
        nboxes += 1
        i += stride
        if time.perf_counter() - t0 > budget_s:
            break
    return nboxes, cells, busy, time.perf_counter() - t0


def cpu_arm(seed, budget_s, small, offset=0, noflush=False, resident=True):
    """the reference's own five entry points (compiled from the unmodified sources) on all host cores, one process per
    core; `noflush` = the fairness build without dynprog_simd.c's per-column _mm_clflush"""
    import multiprocessing as mp
    import benchgen
    from harness import ref_available
    benchgen.build()                        # before the workers start
    kind = ("reference_noflush" if noflush else "reference") if ref_available(noflush) else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, seed, offset + w, cores, budget_s, small, resident) for w in range(cores)])
    nboxes = sum(r[0] for r in res)
    cells = sum(r[1] for r in res)
    wall = max(r[3] for r in res)
    busy = sum(r[2] for r in res)
    # throughput of the DP calls alone on all cores (box generation / genome set-up of the harness excluded)
    dp_s = busy / cores if busy > 0 else wall
    return {"kind": "reference" if kind.startswith("reference") else "port", "variant": kind, "cores": cores, "boxes": nboxes,
            "cells": cells, "wall_s": dp_s, "harness_wall_s": wall, "gcups": cells / dp_s / 1e9}


def bench_config(args):
    """the workload both arms are run on (the reference arm times a bounded sample of the same generator and seed)"""
    return {"workload": WORKLOAD, "boxes_per_gpu": args.boxes, "seed": args.seed, "modemask": args.modemask,
            "stratum": "production 15-150 bp" if args.small else "BASELINE configs[1] 50-2000 bp",
            "genome": "uploaded with every box (characters + probability arrays)" if args.upload_genome else
                      "resident on the device in the reference's compressed layout; boxes carry coordinates, MaxEnt evaluated on the device "
                      "(synthetic model tables)"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    vals = []
    last = None
    for step in range(args.warmup + args.steps):
        last = cpu_arm(args.seed, args.ref_step_seconds, args.small, offset=step * 100003, resident=not args.upload_genome)
        if step >= args.warmup:
            vals.append(last)
    cells = sum(v["cells"] for v in vals)
    wall = sum(v["wall_s"] for v in vals)
    value = cells / wall / 1e9
    sample = "%d boxes of the same generator and seed per step (%.0f s per step on %d processes); stock reference DP incl. its " \
             "per-column _mm_clflush (SURVEY.md F5); cells counted by the oracle" % (
                 sum(v["boxes"] for v in vals) // max(1, len(vals)), args.ref_step_seconds, last["cores"])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, len(vals)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8/int16 saturating (AVX2)", "data": "synthetic",
            "config": bench_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_cpu_baseline:
        nf = cpu_arm(args.seed, args.ref_step_seconds, args.small, offset=7 * 100003, noflush=True, resident=not args.upload_genome)
        if nf["variant"] == "reference_noflush":
            line["cpu_baseline"]["noflush"] = {"value": nf["gcups"], "unit": UNIT, "cores": nf["cores"],
                                               "sample": "%d boxes, same sources with the _mm_clflush statements of dynprog_simd.c defined away "
                                                         "(oracle/refbuild/noflush.h): identical results" % nf["boxes"]}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# second kernel: stage-2 chaining, reported under the "chain" key
# ------------------------------------------------------------------------------------------------
CHAIN_BYTES_PER_HIT = 28        # 4 B position + 4 B score + 20 B link per (querypos, hit)


CHAIN_SM_COUNT, CHAIN_CLOCK_GHZ, CHAIN_WARP_INSTR_PER_HIT = 148, 1.965, 770.0


def chain_section(eng, args, rank, world, barrier, allreduce, MAX, SUM):
    """`--chain-problems` synthetic chaining problems per GPU (cDNAs of 2-11 exons, 1 % error, against 50-200 kb
    regions, k = 8: tests/chaingen.py), `--chain-distinct` distinct ones tiled.  Device-resident time per step from CUDA
    events around zeroing + kernel; e2e through GmapChain_batch_run from host buffers."""
    import numpy as np
    import chaingen
    import chain_harness
    nd = min(args.chain_distinct, args.chain_problems)
    rng = np.random.default_rng(args.seed + 7919 * rank)
    base = [chaingen.make_problem(rng, glen=int(rng.integers(50000, 200000)), nexons=int(rng.integers(2, 12)), exon_len=(80, 400),
                                  err=0.01, k=8, window="full" if i % 2 else "2000", max_nalignments=10) for i in range(nd)]
    cpu_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # before the big batch is built: a quiet heap
        cpu = chain_harness.RefChain() if chain_harness.have_ref() else chain_harness.OracleChain()
        cpu.paths(base[0])
        t0 = time.perf_counter()
        k = 0
        while k < nd and time.perf_counter() - t0 < args.chain_cpu_seconds:
            cpu.paths(base[k])
            k += 1
        dt = time.perf_counter() - t0
        cpu_line = {"value": k / dt, "unit": "align_compute_lookback calls/s", "cores": 1,
                    "kind": "reference" if chain_harness.have_ref() else "port",
                    "sample": "first %d of the distinct problems, one thread, through ctypes (%.1f s)" % (k, dt)}
    eng.chain_setup(**chain_harness.SETUP)
    b = eng.chain_batch()
    n = 0
    while n < args.chain_problems:
        b.add(base[n % nd])
        n += 1
    b.upload()
    for _ in range(max(args.warmup, 3)):
        b.run_resident()
    launches0 = eng.launch_count()
    barrier()
    dev_ms = sum(b.run_resident() for _ in range(args.steps))
    b.run()                                             # untimed: sizes and page-locks the host-side result buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.run()                                         # H2D of the pools, kernel, D2H of paths
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    launches = eng.launch_count() - launches0
    digest = b.digest()
    hits, h2d, d2h = b.nhits(), b.h2d_bytes(), b.d2h_bytes()
    dev_ms_max, e2e_ms_max = allreduce(dev_ms, MAX), allreduce(e2e_ms, MAX)
    tot_problems, tot_hits = allreduce(n, SUM), allreduce(hits, SUM)
    out = None
    if rank == 0:
        ms = dev_ms_max / args.steps
        peak, peak_src = measured_peaks()
        achieved = hits * CHAIN_BYTES_PER_HIT / (dev_ms / args.steps / 1e3) / 1e9
        out = {"metric": "chain_problems_per_s", "value": tot_problems / (ms / 1e3), "unit": "align_compute_lookback calls/s",
               "hits_per_s": tot_hits / (ms / 1e3), "ms_per_step": ms,
               "config": {"workload": "stage-2 chaining: synthetic spliced cDNAs (2-11 exons of 80-400 nt, 1 % error) against 50-200 kb "
                                      "regions, 8-mer hits, lookback direction, gmap defaults", "problems_per_gpu": n,
                          "distinct_problems": nd, "hits_per_gpu": int(hits), "kernel": "gmapchain_kernel (one warp per problem)"},
               # the kernel is a dependent walk bound by instruction issue, not by bytes (ncu, profiles/r01s2_chain_kernel_metrics.csv:
               # 27.6 G warp instructions for 35.6 M hits = ~770 per hit, issue slots 66.5 % active, DRAM 3.4 GB per launch):
               # the model is hits/s at one warp instruction per scheduler and clock; the HBM figure of SURVEY.md section 8d
               # (hits x 28 B) is kept beside it
               "roofline": {"bound": "issue", "achieved": tot_hits / world / (ms / 1e3) / 1e9,
                            "peak": CHAIN_SM_COUNT * 4 * CHAIN_CLOCK_GHZ / CHAIN_WARP_INSTR_PER_HIT, "unit": "G hits/s",
                            "frac": (tot_hits / world / (ms / 1e3) / 1e9) / (CHAIN_SM_COUNT * 4 * CHAIN_CLOCK_GHZ / CHAIN_WARP_INSTR_PER_HIT),
                            "traffic": None,
                            "peak_source": "model: 148 SMs x 4 schedulers x 1.965 GHz / 770 warp instructions per hit (measured by ncu, "
                                           "profiles/r01s2_chain_kernel_metrics.csv)",
                            "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                                    "note": "algorithmic bytes = hits x 28 B (SURVEY.md section 8d): not the bound"}},
               "e2e": {"value": tot_problems / (e2e_ms_max / args.steps / 1e3), "unit": "align_compute_lookback calls/s",
                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / args.steps},
               "gpu_launches": int(launches), "digest": "%016x" % digest}
        if cpu_line is not None:
            out["cpu_baseline"] = cpu_line
    b.free()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def measure_batch(eng, batch, args, barrier, timed_kinds=True):
    """device-only steps (resident inputs, CUDA events), then the device ABI from host buffers, then the entry-point API
    (device + pair-list replay); returns the per-rank sums in ms"""
    batch.upload()                                      # inputs resident in HBM from here on
    for _ in range(args.warmup):
        batch.run_resident()
    launches0 = eng.launch_count()
    barrier()
    w0 = time.perf_counter()
    dev_ms, full_ms = 0.0, 0.0
    kind_ms = [0.0, 0.0, 0.0, 0.0]
    for _ in range(args.steps):
        dev_ms += batch.run_resident()                  # CUDA events on the launching stream, kernels only
        f_ms, _t = eng.last_kernel_ms()                 # each kernel's own events, on the stream it runs on
        full_ms += f_ms
        kind_ms = [x + y for x, y in zip(kind_ms, eng.last_kernel_ms4())]
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1000.0
    launches = eng.launch_count() - launches0
    batch.download()
    digest = batch.digest()
    d2h = batch.d2h_bytes()
    # device ABI from host buffers: H2D + kernels + D2H (one untimed pass first: streams, events, buffers)
    batch.run_device()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        batch.run_device()
    barrier()
    e2e_dev_ms = (time.perf_counter() - e0) * 1000.0
    assert batch.digest() == digest, "results changed between repetitions"
    # entry-point API: the same plus the replay of every edit script into the pair list the reference returns
    batch.run()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        batch.rewind()
        batch.run()
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1000.0
    return {"dev_ms": dev_ms, "full_ms": full_ms, "kind_ms": kind_ms, "wall_ms": wall_ms, "launches": launches, "digest": digest,
            "d2h": d2h, "h2d": batch.h2d_bytes(), "e2e_dev_ms": e2e_dev_ms, "e2e_ms": e2e_ms}


def program_section(args, world):
    """Whole program on a scaled BASELINE config 3 (synthetic genome + spliced cDNAs, gmap_build index made with the
    reference's own pipeline): gmap.sm100 on `world` GPUs beside the stock gmap.avx2 on all host cores, both printing -A
    alignments with -O; reports cDNAs/s by wall clock and by gmap's own "Processed ..." line.  Rank 0 only."""
    import re
    import tempfile
    import program_cases
    import progdata
    have = all(os.path.exists(x) for x in (progdata.AVX2, progdata.SM100, os.path.join(progdata.REFBIN, "gmapindex")))
    if not have:
        return {"unavailable": "needs oracle/_ref (gmap.avx2, gmap_build tools) and integration/_build/gmap.sm100"}
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    out = {"workload": "scaled BASELINE configs[2]: %d synthetic spliced cDNAs (1-3 kb, 2-12 exons, 1 %% error) against a %d Mb synthetic "
                       "genome, gmap_build index (k = 15), gmap -A -O" % (args.program_cdnas, args.program_genome_mb),
           "cores": cores, "n_gpus": world}
    with tempfile.TemporaryDirectory() as td:
        t0 = time.time()
        genome, cdna = progdata.generate(td, "config3", args.program_genome_mb * 1000000, 4, args.program_cdnas, seed=args.seed)
        progdata.build_index(genome, os.path.join(td, "db"), "syn")
        out["setup_s"] = round(time.time() - t0, 1)
        case = ["-D", os.path.join(td, "db"), "-d", "syn", "-A", cdna]

        def timed(exe, threads, env=None):
            t1 = time.time()
            text, err = program_cases.run(exe, case, threads=threads, env=env)
            dt = time.time() - t1
            m = re.search(r"Processed (\d+) queries in ([0-9.]+) seconds", err)
            own = (int(m.group(1)) / float(m.group(2))) if m and float(m.group(2)) > 0 else None
            return text, err, dt, own
        timed(progdata.AVX2, cores)                                  # page cache
        want, _, dt_ref, own_ref = timed(progdata.AVX2, cores)
        out["reference"] = {"exe": "gmap.avx2 (stock, incl. _mm_clflush)", "threads": cores, "cdnas_per_s": args.program_cdnas / dt_ref,
                            "cdnas_per_s_gmap_line": own_ref, "seconds": dt_ref}
        nf = progdata.AVX2 + "_noflush"
        if os.path.exists(nf):
            text, _, dt_nf, own_nf = timed(nf, cores)
            out["reference_noflush"] = {"exe": "gmap.avx2 built without dynprog_simd.c's _mm_clflush (SURVEY.md F5)", "threads": cores,
                                        "cdnas_per_s": args.program_cdnas / dt_nf, "cdnas_per_s_gmap_line": own_nf,
                                        "identical_output": text == want}
        env = dict(os.environ, GMAP_SM100_STATS="1", GMAP_SM100_DEVICES=",".join(str(k) for k in range(world)))
        best = None
        for threads in [int(x) for x in args.program_threads.split(",")]:
            got, err, dt, own = timed(progdata.SM100, threads, env)
            row = {"threads": threads, "cdnas_per_s": args.program_cdnas / dt, "cdnas_per_s_gmap_line": own, "seconds": dt,
                   "identical_output": got == want,
                   "runtime": [l for l in err.splitlines() if l.startswith("gmap.sm100") and ("runtime" in l or "DP calls" in l or "stage 2:" in l)]}
            out.setdefault("sm100_runs", []).append(row)
            if row["identical_output"] and (best is None or row["cdnas_per_s"] > best["cdnas_per_s"]):
                best = row
        if best is not None:
            out["cdnas_per_s"] = best["cdnas_per_s"]
            out["cdnas_per_s_gmap_line"] = best["cdnas_per_s_gmap_line"]
            out["threads"] = best["threads"]
            out["identical_output"] = True
            out["ratio_to_reference"] = best["cdnas_per_s"] / out["reference"]["cdnas_per_s"]
        else:
            out["identical_output"] = False
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import benchgen
    from gmap_2024_b200 import Engine

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # the ranks of one node share its host cores: each rank's replay (the host half of `e2e`) gets its share of them instead
    # of one thread per core per rank (GMAPDP_REPLAY_THREADS is the library's own knob; default: all cores)
    cores_all = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if world > 1 and "GMAPDP_REPLAY_THREADS" not in os.environ:
        os.environ["GMAPDP_REPLAY_THREADS"] = str(max(2, cores_all // env_int("LOCAL_WORLD_SIZE", world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the DP engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    import torch.distributed as d2
    MAX, SUM = (d2.ReduceOp.MAX, d2.ReduceOp.SUM) if world > 1 else (None, None)
    from gmap_2024_b200.sharding import shard_range

    eng = Engine(local)
    batch = eng.batch(2000, 2030)
    t0 = time.time()
    i0, i1 = shard_range(rank, world, args.boxes)
    resident = not args.upload_genome
    genome_bytes = 0
    if resident:                            # the rank's genome: laid out while the boxes are queued, uploaded once (untimed)
        benchgen.resident_begin(batch, args.seed, min(4000000000, (i1 - i0) * 2600 + 1000000))
    if benchgen.fill_batch(batch, args.seed, i0, i1 - i0, 1, args.small, args.modemask) < 0:    # untimed: synthetic input
        raise SystemExit("bench.py: the synthetic genome outgrew its buffer")
    if resident:
        genome_bytes = benchgen.resident_attach(eng)
    gen_s = time.time() - t0
    cells, cells8 = batch.cells(), batch.cells8()

    sampler = ClockSampler(local)
    sampler.start()
    m = measure_batch(eng, batch, args, barrier)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    dev_ms_max = allreduce(m["dev_ms"], MAX)
    wall_ms_max = allreduce(m["wall_ms"], MAX)
    e2e_ms_max = allreduce(m["e2e_ms"], MAX)
    e2e_dev_ms_max = allreduce(m["e2e_dev_ms"], MAX)
    tot_cells = allreduce(cells, SUM)
    tot_launches = allreduce(m["launches"], SUM)
    tot_boxes = allreduce(batch.nboxes(), SUM)
    tot_calls = allreduce(batch.ncalls(), SUM)
    cf, cf8 = batch.cells_full()
    info = eng.device_info()
    batch.free()

    # second stratum: production-size boxes (15-150 bp per side: > 95 % of GMAP's own calls)
    stratum = None
    if args.stratum_boxes > 0 and not args.small:
        sb = eng.batch(2000, 2030)
        j0, j1 = shard_range(rank, world, args.stratum_boxes)
        if resident:
            benchgen.resident_begin(sb, args.seed, (j1 - j0) * 400 + 1000000)
        if benchgen.fill_batch(sb, args.seed, j0, j1 - j0, 1, True, args.modemask) < 0:
            raise SystemExit("bench.py: the synthetic genome outgrew its buffer")
        if resident:
            benchgen.resident_attach(eng)
        s_cells, s_calls, s_boxes = sb.cells(), sb.ncalls(), sb.nboxes()
        sm = measure_batch(eng, sb, args, barrier)
        s_dev, s_e2e, s_e2e_dev = allreduce(sm["dev_ms"], MAX), allreduce(sm["e2e_ms"], MAX), allreduce(sm["e2e_dev_ms"], MAX)
        s_tot_cells, s_tot_calls = allreduce(s_cells, SUM), allreduce(s_calls, SUM)
        tot_launches += allreduce(sm["launches"], SUM)
        if rank == 0:
            stratum = {"config": {"workload": "the same generator at 15-150 bp per side (production-size boxes)", "boxes_per_gpu": args.stratum_boxes,
                                  "seed": args.seed, "calls": int(s_tot_calls), "device_boxes_rank0": int(s_boxes), "cells_per_step": int(s_tot_cells)},
                       "value": s_tot_cells / (s_dev / args.steps / 1e3) / 1e9, "unit": UNIT, "ms_per_step": s_dev / args.steps,
                       "calls_per_s": s_tot_calls / (s_dev / args.steps / 1e3),
                       "e2e": {"value": s_tot_cells / (s_e2e / args.steps / 1e3) / 1e9, "unit": UNIT, "ms_per_step": s_e2e / args.steps,
                               "calls_per_s": s_tot_calls / (s_e2e / args.steps / 1e3),
                               "h2d_bytes_per_step": int(sm["h2d"]), "d2h_bytes_per_step": int(sm["d2h"])},
                       "e2e_device": {"value": s_tot_cells / (s_e2e_dev / args.steps / 1e3) / 1e9, "unit": UNIT, "ms_per_step": s_e2e_dev / args.steps,
                                      "calls_per_s": s_tot_calls / (s_e2e_dev / args.steps / 1e3)},
                       "kernel_ms": {"single": sm["kind_ms"][0] / args.steps, "end": sm["kind_ms"][1] / args.steps,
                                     "genome": sm["kind_ms"][2] / args.steps, "cdna": sm["kind_ms"][3] / args.steps}}
        sb.free()

    # third stratum: the input decorations of SURVEY.md section 8d (N in the genome, lower-case / IUPAC query characters,
    # tandem repeats) on the same generator
    decorated = None
    if args.decorated_boxes > 0 and not args.small:
        db = eng.batch(2000, 2030)
        j0, j1 = shard_range(rank, world, args.decorated_boxes)
        if resident:
            benchgen.resident_begin(db, args.seed, (j1 - j0) * 2600 + 1000000)
        if benchgen.fill_batch(db, args.seed, j0, j1 - j0, 1, False, args.modemask, decor=True) < 0:
            raise SystemExit("bench.py: the synthetic genome outgrew its buffer")
        if resident:
            benchgen.resident_attach(eng)
        d_cells = db.cells()
        dm = measure_batch(eng, db, args, barrier)
        d_dev, d_e2e = allreduce(dm["dev_ms"], MAX), allreduce(dm["e2e_ms"], MAX)
        d_tot = allreduce(d_cells, SUM)
        tot_launches += allreduce(dm["launches"], SUM)
        if rank == 0:
            decorated = {"config": {"workload": "the same generator with 0.3 % N in the genomic segments, 0.2 % lower-case and 0.1 % IUPAC query "
                                                "characters, 5 % of the single gaps tandem repeats (SURVEY.md section 8d)",
                                    "boxes_per_gpu": args.decorated_boxes, "cells_per_step": int(d_tot)},
                         "value": d_tot / (d_dev / args.steps / 1e3) / 1e9, "unit": UNIT, "ms_per_step": d_dev / args.steps,
                         "e2e": {"value": d_tot / (d_e2e / args.steps / 1e3) / 1e9, "unit": UNIT, "ms_per_step": d_e2e / args.steps},
                         "digest": "%016x" % dm["digest"]}
        db.free()

    chain = None
    if args.chain_problems > 0:
        chain = chain_section(eng, args, rank, world, barrier, allreduce, MAX, SUM)
    eng.close()

    program = None
    if args.program_cdnas > 0:
        barrier()
        if rank == 0:
            program = program_section(args, world)
        barrier()

    if rank == 0:
        ms_per_step = dev_ms_max / args.steps
        value = tot_cells / (ms_per_step / 1e3) / 1e9
        e2e_value = tot_cells / (e2e_ms_max / args.steps / 1e3) / 1e9
        e2e_dev_value = tot_cells / (e2e_dev_ms_max / args.steps / 1e3) / 1e9
        hbm_peak, hbm_src = measured_peaks()
        ipeak, ipeak_src = int_peak()
        # roofline of the dominant kernel of a step, gmapdp_dp_kernel<0> (the full fills of the single-gap boxes), from THIS
        # rank's launch and that kernel's own average duration.  Bound: integer issue (18 algorithmic
        # ops per full cell); the HBM figures are reported beside it -- the kernel moves far fewer bytes than the
        # "algorithmic" 2-3 B/cell (scores never leave the SM), so HBM is not what bounds it.
        full_ms = m["full_ms"]
        dom_ms = (full_ms / args.steps) if full_ms > 0 else (m["dev_ms"] / args.steps)
        dom_cells = cf if full_ms > 0 else cells
        dom_ops = OPS_PER_CELL_FULL if full_ms > 0 else OPS_PER_CELL_TRI
        algo_bytes = (cf8 * BYTES_PER_CELL_8 + (cf - cf8) * BYTES_PER_CELL_16) if full_ms > 0 else \
            (cells8 * BYTES_PER_CELL_8 + (cells - cells8) * BYTES_PER_CELL_16)
        achieved_int = dom_cells * dom_ops / (dom_ms / 1e3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        # the committed ncu figure is for the default launch (1M boxes, all modes); other sizes report null
        if os.path.exists(tp) and args.boxes == 1000000 and args.modemask == 31 and not args.small:
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        hbm = {"algorithmic_bytes_per_launch": int(algo_bytes), "algorithmic_gbs": algo_bytes / (dom_ms / 1e3) / 1e9, "peak": hbm_peak,
               "unit": "GB/s", "peak_source": hbm_src}
        hbm["frac_algorithmic"] = hbm["algorithmic_gbs"] / hbm_peak
        if traffic:
            hbm["measured_gbs"] = traffic / (dom_ms / 1e3) / 1e9
            hbm["frac_measured"] = hbm["measured_gbs"] / hbm_peak
        clocks = sampler.summary()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8/int16 saturating (int32 lanes)", "data": "synthetic",
                "config": bench_config(args),
                "run": {"calls": int(tot_calls), "device_boxes": int(tot_boxes), "cells_per_step": int(tot_cells), "l2": "inputs_exceed_l2" if not args.small else "small",
                        "grid_blocks": info["grid_blocks"], "block_threads": info["block_threads"], "parallelism": "shard%d" % world,
                        "host_cores": cores_all, "replay_threads_per_rank": os.environ.get("GMAPDP_REPLAY_THREADS", "all cores"),
                        "wall_ms_per_step": wall_ms_max / args.steps, "input_generation_s": gen_s,
                        "resident_genome_bytes": int(genome_bytes)},
                "roofline": {"bound": "int", "achieved": achieved_int, "peak": ipeak, "unit": "Tera int ops/s", "frac": achieved_int / ipeak,
                             "traffic": traffic, "peak_source": ipeak_src, "ops_per_cell": dom_ops,
                             "kernel": "gmapdp_dp_kernel<0> (single gaps: full fills)" if full_ms > 0 else "all kernels",
                             "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / (m["dev_ms"] / args.steps), "kernel_cells": int(dom_cells),
                             "hbm": hbm,
                             "other_kernels": {"names": ["gmapdp_dp_kernel<1> (end gaps)", "gmapdp_dp_kernel<2> (genome gaps)",
                                                         "gmapdp_dp_kernel<3> (cdna gaps)"],
                                               "ms": [x / args.steps for x in m["kind_ms"][1:]],
                                               "note": "E-only fills, searches and bridges; the four kernels of a step run back to back "
                                                       "on one stream, each bracketed by its own CUDA events"}},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(m["h2d"]), "d2h_bytes_per_step": int(m["d2h"]),
                        "ms_per_step": e2e_ms_max / args.steps,
                        "path": "GmapDP_batch_run: H2D, kernels, D2H and the replay of every edit script into pair lists "
                                "(host threads: GMAPDP_REPLAY_THREADS, default all)"},
                "e2e_device": {"value": e2e_dev_value, "unit": UNIT, "ms_per_step": e2e_dev_ms_max / args.steps,
                               "path": "gmapdp_run_batch (device ABI): H2D, kernels, D2H of results and scripts, no pair lists"},
                "gpu_launches": int(tot_launches), "clocks": clocks, "digest": "%016x" % m["digest"]}
        if stratum is not None:
            line["strata"] = {"production": stratum}
        if decorated is not None:
            line.setdefault("strata", {})["decorated"] = decorated
        if chain is not None:
            line["chain"] = chain
            line["gpu_launches"] += chain["gpu_launches"]
        if program is not None:
            line["program"] = program
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_arm(args.seed, args.cpu_seconds, args.small, resident=not args.upload_genome)
            line["cpu_baseline"] = {"value": c["gcups"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
                                    "sample": "first %d boxes of the same generator and seed, %.0f s on %d processes%s; cells counted by the oracle" % (
                                        c["boxes"], args.cpu_seconds, c["cores"],
                                        " (stock reference DP incl. its per-column _mm_clflush, SURVEY.md F5)" if c["kind"] == "reference" else "")}
            nf = cpu_arm(args.seed, args.cpu_seconds / 2, args.small, noflush=True, resident=not args.upload_genome)
            if nf["variant"] == "reference_noflush":
                line["cpu_baseline"]["noflush"] = {"value": nf["gcups"], "unit": UNIT, "cores": nf["cores"],
                                                   "sample": "%d boxes; same sources with the _mm_clflush statements of dynprog_simd.c defined away "
                                                             "(oracle/refbuild/noflush.h), identical results" % nf["boxes"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--boxes", type=int, default=1000000, help="boxes per GPU (BASELINE.json configs[1]: 1M)")
    ap.add_argument("--seed", type=int, default=20241018)
    ap.add_argument("--upload-genome", action="store_true", help="send genomic characters and probability arrays with every box instead of "
                    "coordinates into a genome resident on the device (the round-1 form of the workload)")
    ap.add_argument("--small", action="store_true", help="run the main measurement on the production-size stratum (15-150 bp) instead")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-step-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chain-problems", type=int, default=32768, help="stage-2 chaining problems per GPU for the \"chain\" key (0 = skip)")
    ap.add_argument("--chain-distinct", type=int, default=256)
    ap.add_argument("--chain-cpu-seconds", type=float, default=5.0)
    ap.add_argument("--stratum-boxes", type=int, default=1000000, help="production-size boxes per GPU for strata.production (0 = skip)")
    ap.add_argument("--decorated-boxes", type=int, default=200000, help="boxes per GPU for strata.decorated (0 = skip)")
    ap.add_argument("--program-cdnas", type=int, default=10000, help="whole-program leg: synthetic cDNAs (0 = skip)")
    ap.add_argument("--program-genome-mb", type=int, default=20)
    ap.add_argument("--program-threads", default="128,256", help="gmap.sm100 -t values to try (the best identical run is reported)")
    ap.add_argument("--modemask", type=int, default=31, help="diagnostics: bit k keeps mode k (single,genome,cdna,end5,end3); 31 = the benchmark config")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
