#!/usr/bin/env python
"""bench.py -- RNAcode scoring hot path on B200: codon-DP cells/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload genomic|short|wide]

A "step" is one pass of the hot path (pack -> sigma -> DP -> HSS digest replay) over one batch of
synthetic alignment blocks with their null alignments.  Default workload: the block shapes of the
reference's examples/genomic.maf (BASELINE.json configs[1]) filled by the seeded generator of
SURVEY.md 8(d), n = 1000 null alignments per block.

value   : DP cells/s with inputs already resident in HBM (rc_batch_run only), CUDA events, max over ranks.
e2e     : same metric through the C ABI from host (pinned) buffers: rc_batch_create + upload (H2D) + run +
          download (D2H) + destroy inside the timed region.
roofline: the DP kernel (k_dp) against the FP32/ALU issue ceiling; achieved = algorithmic
          FP32 lane-ops (6 per DP cell) / k_dp time from CUDA events on its own stream.
N > 1   : one process per GPU (torchrun), every rank scores its own blocks (weak scaling, no collective on
          the data path); barrier + max over ranks for the timing only.
--impl reference : the reference's own CPU implementation (oracle/_ref/RNAcode_ref built from the unmodified
          sources, else the oracle port) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rnacode_b200 import synth  # noqa: E402

GENOMIC_SHAPES = [(9, 319), (10, 4806), (8, 3), (6, 2), (6, 312), (8, 342), (4, 76), (7, 97), (4, 143), (4, 86), (6, 84)]
WORKLOADS = {
    # name: (list of (N, cols), n_samples, seed, description)
    "genomic": (GENOMIC_SHAPES, 1000, 2, "synthetic MAF with the 11 block shapes of examples/genomic.maf, -n 1000"),
    "short": ([(10, 120)] * 2000, 100, 1, "synthetic MAF 2000 blocks x 10 species x 120 cols, -n 100 (config 3 shape)"),
    "wide": ([(50, 5000)] * 1, 200, 3, "synthetic MAF 1 block x 50 species x 5000 cols, -n 200 (config 4 shape, reduced n)"),
    "hundred": ([(100, 1000)] * 4, 250, 4, "synthetic MAF 4 blocks x 100 species x 1000 cols, -n 250 (config 5 row count)"),
    "hundred_short": ([(100, 200)] * 200, 100, 5, "synthetic MAF 200 blocks x 100 species x 200 cols, -n 100 (config 5 row count, short blocks)"),
    "mid_short": ([(30, 200)] * 500, 100, 6, "synthetic MAF 500 blocks x 30 species x 200 cols, -n 100"),
    "mid": ([(10, 800)] * 16, 1000, 7, "synthetic MAF 16 blocks x 10 species x 800 cols, -n 1000"),
    "mid12": ([(10, 1200)] * 8, 1000, 9, "synthetic MAF 8 blocks x 10 species x 1200 cols, -n 1000"),
    "mid5": ([(10, 500)] * 32, 1000, 10, "synthetic MAF 32 blocks x 10 species x 500 cols, -n 1000"),
    "mid24": ([(10, 2400)] * 4, 1000, 11, "synthetic MAF 4 blocks x 10 species x 2400 cols, -n 1000"),
    "n17": ([(17, 3000)] * 2, 500, 12, "synthetic MAF 2 blocks x 17 species x 3000 cols, -n 500"),
    "n14": ([(14, 3000)] * 2, 500, 13, "synthetic MAF 2 blocks x 14 species x 3000 cols, -n 500"),
    "n17s": ([(17, 150)] * 500, 100, 14, "synthetic MAF 500 blocks x 17 species x 150 cols, -n 100"),
    "mid_wide": ([(50, 800)] * 4, 250, 8, "synthetic MAF 4 blocks x 50 species x 800 cols, -n 250"),
    # the default workload with the frameshift density of the real examples/genomic.maf (3 % of the codon pairs of its
    # 10 x 4806 block carry a frameshift of some species; SURVEY 8(d)'s generator gives 36 %) and without any gap
    "genomic_lowgap": (GENOMIC_SHAPES, 1000, 2, "genomic.maf block shapes, gap rate 0.0005 (frameshift density of the real file), -n 1000", 0.0005),
    "genomic_gapfree": (GENOMIC_SHAPES, 1000, 2, "genomic.maf block shapes, gap-free (the 6-op cell of SURVEY 8(d)), -n 1000", 0.0),
}
METRIC = "codon_dp_cells_per_s"
UNIT = "cells/s"
OPS_PER_CELL = 6.0  # 3 FADD + MAX3 (2 FMNMX) + 1 FADD for a cell without frameshift


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def build_workload(name, rank):
    shapes, n, seed, desc = WORKLOADS[name][:4]
    gap_rate = WORKLOADS[name][4] if len(WORKLOADS[name]) > 4 else 0.0067
    blocks = []
    for i, (N, cols) in enumerate(shapes):
        idx = rank * 100000 + i
        rows = synth.synth_block(seed, idx, N, cols, gap_rate=gap_rate)
        sf, sr = synth.synth_scores(seed, idx, N)
        blocks.append((rows, sf, sr, idx))
    return blocks, n, seed, desc


def workload_cells(blocks, n):
    return float(sum(synth.cells(r.shape[0], synth.ungapped_len(r), n) for r, _, _, _ in blocks))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        # NVML in-process (about a millisecond per sample, so even a 0.3 s timed region gets dozens of samples); the
        # nvidia-smi command line of the recipe is the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx), ""] + [("Active" if mask & b else "Not Active") for _, b in bits])
                self.stop_flag.wait(0.01)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ------------------------------------------------------------------------------------------------
def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "RNAcode_ref")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def cpu_sample_blocks(workload, target_cells):
    """Bounded sample of the workload: every block shape, with n_sub null alignments chosen so that the
    sample holds about target_cells DP cells."""
    blocks, n, seed, desc = build_workload(workload, 0)
    per_aln = sum(synth.cells(r.shape[0], synth.ungapped_len(r), 0) for r, _, _, _ in blocks)
    reps = 1
    n_sub = int(round(target_cells / per_aln)) - 1
    if n_sub > n:
        n_sub = n
    if n_sub < 1:
        n_sub = 1
    return blocks, n_sub, per_aln, reps


def run_reference_cli(workload, cores, target_cells, tmpdir):
    """One process of the unmodified reference per block (it is single-threaded), `cores` at a time.
    Returns (cells, seconds, description)."""
    exe = ref_binary()
    blocks, n_sub, per_aln, _ = cpu_sample_blocks(workload, target_cells)
    # memory guard: the reference allocates 3*N*3*(L+1)^2*4 bytes per block (src/misc.c:33-55)
    jobs = []
    for i, (rows, _, _, _) in enumerate(blocks):
        L = synth.ungapped_len(rows)
        if rows.shape[0] <= 2 or L < 3:
            continue
        p = os.path.join(tmpdir, "blk%d.maf" % i)
        if not os.path.exists(p):
            synth.to_maf([rows], p)
        mem = 3.0 * rows.shape[0] * 3 * (L + 1) ** 2 * 4
        jobs.append((mem, p, synth.cells(rows.shape[0], L, n_sub)))
    # replicate the job list so that every core has work: the reference cannot split a block
    small = [j for j in jobs if j[0] < 2e9]
    big = [j for j in jobs if j[0] >= 2e9]
    try:
        avail = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") * 0.5
    except Exception:
        avail = 32e9
    big_copies = max(1, min(cores, int(avail // max(j[0] for j in big)))) if big else 0
    joblist = []
    for c in range(big_copies):
        joblist += big
    while len(joblist) < cores and small:
        joblist += small
    if not joblist:
        joblist = jobs
    t0 = time.perf_counter()
    running = []
    queue = list(joblist)
    total_cells = 0.0
    while queue or running:
        while queue and len(running) < cores:
            mem, p, c = queue.pop(0)
            pr = subprocess.Popen([exe, "-n", str(n_sub), "--tabular", p], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            running.append((pr, c))
        still = []
        for pr, c in running:
            if pr.poll() is None:
                still.append((pr, c))
            else:
                total_cells += c
        running = still
        if running:
            time.sleep(0.005)
    dt = time.perf_counter() - t0
    desc = "%d reference processes (one per block, %d at a time): workload block shapes at -n %d, big blocks x%d" % (
        len(joblist), cores, n_sub, big_copies)
    return total_cells, dt, desc


def run_oracle_port(workload, target_cells):
    """Single-threaded oracle port (oracle/liboracle.so) on a bounded sample. Returns (cells, seconds, desc)."""
    from tests import oracle_py as op
    orc = op.load()
    blocks, n_sub, per_aln, _ = cpu_sample_blocks(workload, target_cells)
    prm = orc.params()
    t0 = time.perf_counter()
    cells = 0.0
    for rows, sf, sr, idx in blocks:
        L = synth.ungapped_len(rows)
        if L < 3:
            continue
        orc.score_aln(rows, sf, sr, prm)
        smp = synth.synth_samples(7, idx, n_sub, rows.shape[0], rows.shape[1])
        orc.sample_maxima(rows, smp, sf, sr, prm)
        cells += synth.cells(rows.shape[0], L, n_sub)
    dt = time.perf_counter() - t0
    return cells, dt, "oracle port, workload block shapes with native + %d null alignments each" % n_sub


def reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    times, cells = [], 0.0
    tmpdir = tempfile.mkdtemp(prefix="rc_ref_")
    kind = "reference" if ref_binary() else "port"
    desc = ""
    try:
        # size the sample: about 6 s per step on the reference (measured ~0.05-0.09 Gcells/s/core here)
        target = 4.0e8 if kind == "reference" else 2.0e9
        for it in range(args.warmup + args.steps):
            if kind == "reference":
                c, dt, desc = run_reference_cli(args.workload, cores, target, tmpdir)
            else:
                c, dt, desc = run_oracle_port(args.workload, target)
                cores = 1
            if it >= args.warmup:
                times.append(dt)
                cells = c
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    ms = 1e3 * float(np.mean(times))
    value = cells / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][3], "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from rnacode_b200 import capi

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    blocks_np, n, seed, desc = build_workload(args.workload, rank)
    if args.samples:
        n = args.samples
    cells = workload_cells(blocks_np, n)

    # host buffers in pinned memory (the e2e leg copies from them every step)
    keep = []
    blocks = []
    for rows, sf, sr, idx in blocks_np:
        N, cols = rows.shape
        smp = synth.synth_samples(seed, idx, n, N, cols)
        t = torch.from_numpy(smp).pin_memory()
        r = torch.from_numpy(rows.copy()).pin_memory()
        keep += [t, r]
        blocks.append(capi.Block(r.numpy(), sf, sr, t.numpy()))
    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    prm = capi.make_params()
    blosum = np.array(BLOSUM62, dtype=np.int32)

    # --- device-resident leg -------------------------------------------------------------------
    bt = ctx.batch(blocks, prm, blosum)
    bt.upload()
    for _ in range(max(args.warmup, 3)):
        bt.run()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dp_ms, stage_ms, launches, dp_launches = 0.0, {"pack": 0.0, "sigma": 0.0, "dp": 0.0, "hss": 0.0}, 0, 0
    pack_kernel_ms = 0.0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        bt.run()
        st = bt.stats()
        dp_ms += st["ms_dp"]
        pack_kernel_ms += st["ms_pack_kernel"]
        for k in stage_ms:
            stage_ms[k] += st["ms_" + k]
        launches += st["launches"]
        dp_launches += st["dp_launches"]
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    sampler.stop_flag.set()
    sampler.join(timeout=5)
    st = bt.stats()
    fallbacks = st["dense_fallbacks"]
    bt.download()
    best_native = [len(bt.native_hss(i)) for i in range(len(blocks))]
    bt.close()

    # --- end-to-end leg through the C ABI with host buffers ----------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        b2 = ctx.batch(blocks, prm, blosum); b2.upload(); b2.run(); b2.download(); b2.close()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h2d = d2h = 0
    for _ in range(e2e_steps):
        b2 = ctx.batch(blocks, prm, blosum)
        b2.upload()
        b2.run()
        b2.download()
        s2 = b2.stats()
        h2d, d2h = s2["h2d_bytes"], s2["d2h_bytes"]
        _ = [b2.max_scores(i) for i in range(len(blocks))]
        b2.close()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    # --- end-to-end with the null alignments drawn on the GPU (kernel d, the CLIs' default): the host supplies
    # the native rows, the score tables, a flattened tree and one seed per sample
    trees = [capi.Tree(*synth.synth_tree(seed, idx, rows.shape[0])) for rows, _, _, idx in blocks_np]
    seeds = [np.arange(1, n + 1, dtype=np.uint32) + 7919 * i for i in range(len(blocks))]
    ev_blocks = [capi.Block(b.rows, b.scores_fwd, b.scores_rev, None, n_samples=n) for b in blocks]

    def evolve_step():
        b3 = ctx.batch(ev_blocks, prm, blosum)
        for i in range(len(ev_blocks)):
            b3.set_evolve(i, trees[i], seeds[i], capi.RC_RNG_MT19937)
        b3.upload(); b3.run(); b3.download()
        s3 = b3.stats()
        _ = [b3.max_scores(i) for i in range(len(ev_blocks))]
        b3.close()
        return s3
    for _ in range(2):
        evolve_step()
    barrier()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record(stream)
    for _ in range(e2e_steps):
        s3 = evolve_step()
    v1.record(stream)
    barrier()
    e2e_evolve_ms = v0.elapsed_time(v1)

    issue_measured = ctx.calibrate_issue()

    # --- reduce over ranks -----------------------------------------------------------------------------
    vals = torch.tensor([total_ms, e2e_ms, dp_ms, e2e_evolve_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([cells, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, dp_ms_max, e2e_evolve_ms = [float(x) for x in vals.tolist()]
    cells_all, launches_all = [float(x) for x in tot.tolist()]

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = cells_all / (ms_per_step * 1e-3)
        e2e_value = cells_all / (e2e_ms / e2e_steps * 1e-3)
        clocks = sampler.summary()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        nominal_peak = sm_count * 128 * sm_max * 1e6  # lane-issues per second
        dp_per_step_s = (dp_ms / args.steps) * 1e-3   # rank-local k_dp time per step
        achieved = cells * OPS_PER_CELL / dp_per_step_s
        nblocks = len(blocks) * world
        traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "blocks_per_gpu": len(blocks), "n_samples": n, "cells_per_step_per_gpu": cells,
                       "l2_policy": "inputs larger than L2 (sigma tiles + row records + class bytes: %d MiB per step)" % (
                           st["device_bytes"] >> 20),
                       "nominal_unit": "cols*6*(n+1) per block: %.4g per step" % sum(
                           6.0 * b.cols * (n + 1) for b in blocks)},
            "blocks_per_s": nblocks / (ms_per_step * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "blocks_per_s": nblocks / (e2e_ms / e2e_steps * 1e-3)},
            "e2e_gpu_evolve": {"value": cells_all / (e2e_evolve_ms / e2e_steps * 1e-3), "unit": UNIT,
                               "h2d_bytes_per_step": int(s3["h2d_bytes"]), "d2h_bytes_per_step": int(s3["d2h_bytes"]),
                               "ms_per_step": e2e_evolve_ms / e2e_steps,
                               "note": "same C-ABI sequence with the null alignments simulated on the GPU (exact MT19937 mode)"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "fp32_issue", "kernel": "k_dp", "achieved": achieved / 1e12, "peak": nominal_peak / 1e12,
                         "unit": "TFLOP/s", "frac": achieved / nominal_peak,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic and not args.samples else None,
                         "traffic_source": traffic["source"] if traffic and not args.samples else None,
                         "peak_source": "nominal %d SMs x 128 lanes x %.0f MHz (no FP32-issue figure in MEASURED_PEAKS.json)" % (
                             sm_count, sm_max),
                         "peak_measured": issue_measured / 1e12,
                         "frac_of_measured": achieved / issue_measured if issue_measured else None,
                         "ops_per_cell": OPS_PER_CELL, "kernel_ms_per_step": dp_ms / args.steps,
                         "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()}},
            # subsystem (a): k_pack against the measured HBM copy bandwidth; algorithmic bytes = 2 B per character
            # (16 B read + 16 B written per 16 characters)
            "roofline_pack": {"bound": "hbm", "kernel": "k_pack", "unit": "GB/s",
                              "achieved": 2.0 * st["pack_chars"] / (pack_kernel_ms / args.steps * 1e-3) / 1e9 if pack_kernel_ms else None,
                              "peak": peaks.get("hbm_gbs"),
                              "frac": (2.0 * st["pack_chars"] / (pack_kernel_ms / args.steps * 1e-3) / 1e9 / peaks["hbm_gbs"])
                              if pack_kernel_ms and peaks.get("hbm_gbs") else None,
                              "kernel_ms_per_step": pack_kernel_ms / args.steps},
            "dense_fallbacks": int(fallbacks),
            "native_hss_total": int(sum(best_native)),
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(workload):
    cores = os.cpu_count() or 1
    tmpdir = tempfile.mkdtemp(prefix="rc_cpu_")
    try:
        if ref_binary():
            c, dt, desc = run_reference_cli(workload, cores, 6.0e8, tmpdir)
            return {"value": c / dt, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc, "seconds": dt}
        c, dt, desc = run_oracle_port(workload, 4.0e9)
        return {"value": c / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc, "seconds": dt}
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


# NCBI BLOSUM62 (A R N D C Q E G H I L K M F P S T W Y V B Z X *): input data of the benchmark
BLOSUM62 = [
    4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0, -2, -1, 0, -4,
    -1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3, -1, 0, -1, -4,
    -2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3, 3, 0, -1, -4,
    -2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3, 4, 1, -1, -4,
    0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4,
    -1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2, 0, 3, -1, -4,
    -1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4,
    0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3, -1, -2, -1, -4,
    -2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3, 0, 0, -1, -4,
    -1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3, -3, -3, -1, -4,
    -1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1, -4, -3, -1, -4,
    -1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2, 0, 1, -1, -4,
    -1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1, -3, -1, -1, -4,
    -2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1, -3, -3, -1, -4,
    -1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2, -2, -1, -2, -4,
    1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2, 0, 0, 0, -4,
    0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0, -1, -1, 0, -4,
    -3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3, -4, -3, -2, -4,
    -2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1, -3, -2, -1, -4,
    0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4, -3, -2, -1, -4,
    -2, -1, 3, 4, -3, 0, 1, -1, 0, -3, -4, 0, -3, -3, -2, 0, -1, -4, -3, -3, 4, 1, -1, -4,
    -1, 0, 0, 1, -3, 3, 4, -2, 0, -3, -3, 1, -1, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4,
    0, -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0, 0, -2, -1, -1, -1, -1, -1, -4,
    -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1,
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="genomic", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override the number of null alignments per block")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
