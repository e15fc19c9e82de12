#!/usr/bin/env python
"""bench.py -- RNAcode scoring hot path on B200: codon-DP cells/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (pack -> sigma -> DP -> HSS digest replay) over one batch of
synthetic alignment blocks with their null alignments.  Default workload: the block shapes of the
reference's examples/genomic.maf (BASELINE.json configs[1]) filled by the seeded generator of
SURVEY.md 8(d), n = 1000 null alignments per block, six differently seeded copies of the 11 shapes per
step so that the timed region of K = 20 steps lasts about 2 s.

value    : DP cells/s with inputs already resident in HBM (rc_batch_run only), CUDA events, max over ranks.
e2e      : same metric through the C ABI from host (pinned) buffers: rc_batch_create + upload (H2D) + run +
           download (D2H) + destroy inside the timed region.
roofline : the DP kernels against the FP32/ALU issue ceiling; achieved = algorithmic
           FP32 lane-ops (6 per DP cell) / DP kernel time from CUDA events on its own stream.
workloads: (N = 1) the other BASELINE.json configs -- config 3 (10 000 short blocks), config 4's block shape at
           -n 1000, config 5's 100-way rows -- each with cells/s, ms per step, per-stage ms, roofline fraction and e2e.
cli      : (N = 1) wall time of the drop-in CLI RNAcode_b200 (host stages included) on examples/genomic.maf
           --gtf --best-only -n 1000 and on config 3's synthetic MAF: MAF blocks/s end to end (SURVEY 8d metric 2).
sharded  : strong scaling of two fixed global block lists (config 4: 50 x 5000 at -n 1000; config 5: 100-way blocks of
           mixed length, -p 0.05 --stop-early) through the product's sharder (rnacode_b200/shard.py, the mirror of
           integration/rnacode_pipeline.c: cost-weighted LPT over (block, sample-range) units, host gather in input
           order inside the timed region); rank 0 re-scores the whole list alone outside the timed region and compares
           a digest of all HSS + maxima ("shard_parity").
N > 1    : one process per GPU (torchrun).  `value` stays the weak-scaling figure (every rank scores its own copy of the
           default workload, no collective on the data path; barrier + max over ranks for the timing only); the
           strong-scaling figures are in `sharded`.
--impl reference : the reference's own CPU implementation (oracle/_ref/RNAcode_ref built from the unmodified
           sources, else the oracle port) on the box's host cores, on a bounded sample of the same workload
           extrapolated to the workload's own -n (see reference_arm).
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

from rnacode_b200 import shard, synth  # noqa: E402

GENOMIC_SHAPES = [(9, 319), (10, 4806), (8, 3), (6, 2), (6, 312), (8, 342), (4, 76), (7, 97), (4, 143), (4, 86), (6, 84)]
GENOMIC_COPIES = 6
WORKLOADS = {
    # name: (list of (N, cols), n_samples, seed, description[, gap rate])
    "genomic": (GENOMIC_SHAPES * GENOMIC_COPIES, 1000, 2,
                "synthetic MAF with the 11 block shapes of examples/genomic.maf x %d differently seeded copies, -n 1000" % GENOMIC_COPIES),
    "genomic1": (GENOMIC_SHAPES, 1000, 2, "synthetic MAF with the 11 block shapes of examples/genomic.maf, -n 1000"),
    "short": ([(10, 120)] * 10000, 100, 1, "synthetic MAF 10000 blocks x 10 species x 120 cols, -n 100 (config 3)"),
    "short2k": ([(10, 120)] * 2000, 100, 1, "synthetic MAF 2000 blocks x 10 species x 120 cols, -n 100 (config 3 shape)"),
    "wide": ([(50, 5000)] * 1, 200, 3, "synthetic MAF 1 block x 50 species x 5000 cols, -n 200 (config 4 shape, reduced n)"),
    "wide_n1000": ([(50, 5000)] * 2, 1000, 3, "synthetic MAF 2 blocks x 50 species x 5000 cols, -n 1000 (config 4 block shape and n)"),
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
    # the default shapes with the frameshift density of the real examples/genomic.maf (3 % of the codon pairs of its
    # 10 x 4806 block carry a frameshift of some species; SURVEY 8(d)'s generator gives 36 %) and without any gap
    "genomic_lowgap": (GENOMIC_SHAPES, 1000, 2, "genomic.maf block shapes, gap rate 0.0005 (frameshift density of the real file), -n 1000", 0.0005),
    "genomic_gapfree": (GENOMIC_SHAPES, 1000, 2, "genomic.maf block shapes, gap-free (the 6-op cell of SURVEY 8(d)), -n 1000", 0.0),
}
SIDE_WORKLOADS = ["short", "wide_n1000", "hundred", "hundred_short", "genomic_lowgap"]  # reported under "workloads" at N = 1
METRIC = "codon_dp_cells_per_s"
UNIT = "cells/s"
OPS_PER_CELL = 6.0    # 3 FADD + MAX3 (2 FMNMX) + 1 FADD for a cell without frameshift
OPS_PER_CELL_FS = 12.0  # ... and for a cell of a species with a frameshift at that codon


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


def workload_config(name, n_override=0):
    """The `config` object of the JSON line: a pure function of the workload, identical in both arms."""
    shapes, n, seed, desc = WORKLOADS[name][:4]
    n = n_override or n
    blocks, _, _, _ = build_workload(name, 0)
    return {"workload": desc if not n_override else desc + " [-n overridden: %d]" % n, "blocks_per_gpu": len(shapes),
            "n_samples": n, "cells_per_step_per_gpu": workload_cells(blocks, n),
            "nominal_unit": "cols*6*(n+1) per block: %.4g per step" % sum(6.0 * c * (n + 1) for _, c in shapes),
            "l2_policy": "inputs larger than L2 (class bytes + sigma tiles + row records of one step exceed the 126 MB L2 many times over)"}


def frameshift_fraction(blocks_np):
    """g of SURVEY 8(d): fraction of (species, codon) pairs with z != 0 on the forward strand, over the blocks of a
    workload, weighted by the DP cells that read them (a codon at site j of a frame is read by j+1 rows)."""
    num = den = 0.0
    for rows, _, _, _ in blocks_np:
        N, cols = rows.shape
        gap = rows == synth.GAP
        ref_cols = np.nonzero(~gap[0])[0]
        L = len(ref_cols)
        if L < 3:
            continue
        cg = np.concatenate([np.zeros((N, 1), dtype=np.int64), np.cumsum(gap, axis=1)], axis=1)  # gaps before column c
        for f in range(3):
            sites = (L - f) // 3
            if sites <= 0:
                continue
            x = 3 * np.arange(sites) + 3 + f           # 1-based reference position of the codon's last nucleotide
            hi = ref_cols[x - 1] + 1                   # one past the codon's last column
            lo = np.where(x > 3, ref_cols[np.maximum(x - 4, 0)] + 1, 0)
            gk = cg[1:, hi] - cg[1:, lo]               # (N-1, sites) gaps of species k in the block of columns
            g0 = (hi - lo) - 3
            z = (np.abs(gk - g0[None, :]) % 3) != 0
            w = (np.arange(sites) + 1.0)[None, :]
            num += float((z * w).sum())
            den += float(w.sum() * (N - 1))
    return num / den if den else 0.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        # NVML in-process (about a millisecond per sample); the nvidia-smi command line of the recipe is the fallback
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


REF_BIG_BYTES = 2e9     # a block whose Sk arrays exceed this is a "big" block (3*N*3*(L+1)^2*4 bytes, src/misc.c:33-55)
REF_BIG_N = (1, 4)      # the two -n values the big block is timed at (per-alignment time = slope, fixed costs = intercept)


def run_reference_cli(workload, cores, tmpdir, small_budget_s=6.0, with_big=True):
    """The unmodified reference (single-threaded) as one process per block, on a bounded sample of the workload that is
    extrapolated to the workload's own -n:

    * every distinct block shape of the workload once (copies of a shape behave alike);
    * small blocks at the workload's own -n when that fits `small_budget_s` of one core (measured rate otherwise at a
      reduced -n, stated);
    * a block whose Sk arrays need gigabytes (10 x 4806: 8.3 GB, 0.74 s per alignment, 744 s at -n 1000) is timed at
      -n 1 and -n 4 in the same step; the time of one alignment is the slope, the native-only costs (PhyML, copySk) the
      intercept, and T(n) = intercept + (n+1) * slope is the time at the workload's n.  At most cores/2 such processes run
      at once (more are memory-bandwidth starved), the other cores work on the small blocks meanwhile.

    Box throughput = sum over the concurrently running processes of cells(n) / T(n): what the box delivers when it is kept
    full of such blocks.  Returns (value cells/s, wall seconds of the step, description, detail dict)."""
    exe = ref_binary()
    blocks, n, seed, desc = build_workload(workload, 0)
    seen, jobs = set(), []
    for i, (rows, _, _, _) in enumerate(blocks):
        L = synth.ungapped_len(rows)
        if rows.shape[0] <= 2 or L < 3 or rows.shape in seen:
            continue
        seen.add(rows.shape)
        p = os.path.join(tmpdir, "blk%d.maf" % i)
        if not os.path.exists(p):
            synth.to_maf([rows], p)
        mem = 3.0 * rows.shape[0] * 3 * (L + 1) ** 2 * 4
        jobs.append({"path": p, "N": rows.shape[0], "L": L, "mem": mem, "cells1": synth.cells(rows.shape[0], L, 0)})
    small = [j for j in jobs if j["mem"] < REF_BIG_BYTES]
    big = [j for j in jobs if j["mem"] >= REF_BIG_BYTES] if with_big else []
    try:
        avail = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") * 0.6
    except Exception:
        avail = 32e9
    big_slots = 0
    if big:
        big_slots = max(2, min(cores // 2, int(avail // max(j["mem"] for j in big))))
        big_slots -= big_slots % 2  # pairs: one process at each of the two -n values
        big_slots = max(2, big_slots)
    # -n of the small blocks: the workload's own if the whole set fits the budget on the cores left to them
    small_cores = max(1, cores - big_slots)
    est_rate = 6.0e7  # cells/s per core (45-90 Mcells/s)
    n_small = n
    small_cells1 = sum(j["cells1"] for j in small)
    if small and small_cells1 * (n + 1) / est_rate / small_cores > small_budget_s:
        n_small = max(1, int(small_budget_s * small_cores * est_rate / small_cells1) - 1)
    queue = []
    for k in range(big_slots):
        j = big[k // 2 % len(big)]
        queue.append((j, REF_BIG_N[k % 2], True))
    # small blocks: enough copies of the set to keep their cores busy for the duration of the big processes
    reps = 1
    if big and small:
        t_big = 1.5 + (REF_BIG_N[1] + 1) * max(j["cells1"] for j in big) / est_rate
        t_set = small_cells1 * (n_small + 1) / est_rate / small_cores
        reps = max(1, min(64, int(t_big / max(t_set, 1e-3))))
    if not big and small:
        reps = max(1, (cores + len(small) - 1) // len(small))
    for _ in range(reps):
        for j in small:
            queue.append((j, n_small, False))
    t0 = time.perf_counter()
    running, done = [], []
    max_big, max_small = big_slots, small_cores
    nb = ns = 0
    while queue or running:
        k = 0
        while k < len(queue):
            j, nn, is_big = queue[k]
            if (is_big and nb < max_big) or (not is_big and ns < max_small):
                pr = subprocess.Popen([exe, "-n", str(nn), "--tabular", j["path"]], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                running.append((pr, j, nn, is_big, time.perf_counter()))
                nb += is_big
                ns += (not is_big)
                queue.pop(k)
            else:
                k += 1
        still = []
        for pr, j, nn, is_big, ts in running:
            if pr.poll() is None:
                still.append((pr, j, nn, is_big, ts))
            else:
                done.append((j, nn, is_big, time.perf_counter() - ts))
                nb -= is_big
                ns -= (not is_big)
        running = still
        if running:
            time.sleep(0.002)
    wall = time.perf_counter() - t0
    # small blocks: measured rate of the cores that ran them
    small_done = [(j, nn, t) for j, nn, is_big, t in done if not is_big]
    small_rate = 0.0
    if small_done:
        core_seconds = sum(t for _, _, t in small_done)
        small_rate = sum(j["cells1"] * (nn + 1) for j, nn, _ in small_done) / core_seconds  # cells/s of one core
    # big blocks: two-point model per process pair
    big_rate, big_detail = 0.0, None
    if big:
        ta = [t for j, nn, is_big, t in done if is_big and nn == REF_BIG_N[0]]
        tb = [t for j, nn, is_big, t in done if is_big and nn == REF_BIG_N[1]]
        slope = (np.mean(tb) - np.mean(ta)) / (REF_BIG_N[1] - REF_BIG_N[0])
        icpt = np.mean(ta) - (REF_BIG_N[0] + 1) * slope
        j = big[0]
        t_full = icpt + (n + 1) * slope
        big_rate = j["cells1"] * (n + 1) / t_full  # cells/s of one process at the workload's n
        big_detail = {"block": "%dx%d" % (j["N"], j["L"]), "seconds_at_n%d" % REF_BIG_N[0]: float(np.mean(ta)),
                      "seconds_at_n%d" % REF_BIG_N[1]: float(np.mean(tb)), "seconds_per_alignment": float(slope),
                      "native_only_seconds": float(icpt), "extrapolated_seconds_at_n%d" % n: float(t_full),
                      "cells_per_s_per_process": float(big_rate), "concurrent_processes": big_slots}
    value = big_rate * big_slots + small_rate * (small_cores if small_done else 0)
    sample = ("one reference process per block, %d cores: %d processes on the %s block (half at -n %d, half at -n %d; time at -n %d "
              "from the two-point linear model) + the %d small block shapes at -n %d on the other %d cores (x%d)") % (
        cores, big_slots, big_detail["block"] if big_detail else "-", REF_BIG_N[0], REF_BIG_N[1], n, len(small), n_small,
        small_cores, reps) if big else (
        "one reference process per block, %d at a time: the %d block shapes at -n %d (x%d)" % (cores, len(small), n_small, reps))
    detail = {"big": big_detail, "small_cells_per_s_per_core": float(small_rate), "small_n": n_small, "workload_n": n,
              "processes": len(done)}
    return value, wall, sample, detail


def run_oracle_port(workload, target_cells):
    """Single-threaded oracle port (oracle/liboracle.so) on a bounded sample. Returns (cells, seconds, desc)."""
    from tests import oracle_py as op
    orc = op.load()
    blocks, n, seed, desc = build_workload(workload, 0)
    per_aln = sum(synth.cells(r.shape[0], synth.ungapped_len(r), 0) for r, _, _, _ in blocks)
    n_sub = max(1, min(n, int(round(target_cells / per_aln)) - 1))
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
    times, vals = [], []
    tmpdir = tempfile.mkdtemp(prefix="rc_ref_")
    kind = "reference" if ref_binary() else "port"
    desc, detail = "", None
    try:
        for it in range(args.warmup + args.steps):
            if kind == "reference":
                # warm-up steps (untimed; they only page the binary and the inputs in) leave out the gigabyte-sized block
                v, dt, desc, detail = run_reference_cli(args.workload, cores, tmpdir, with_big=it >= args.warmup,
                                                        small_budget_s=6.0 if it >= args.warmup else 1.5)
            else:
                c, dt, desc = run_oracle_port(args.workload, 2.0e9)
                v = c / dt
                cores = 1
            if it >= args.warmup:
                times.append(dt)
                vals.append(v)
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    ms = 1e3 * float(np.mean(times))
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload),
        "sample": desc, "sample_detail": detail,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(workload):
    cores = os.cpu_count() or 1
    tmpdir = tempfile.mkdtemp(prefix="rc_cpu_")
    try:
        if ref_binary():
            v, dt, desc, detail = run_reference_cli(workload, cores, tmpdir)
            return {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc, "seconds": dt, "detail": detail}
        c, dt, desc = run_oracle_port(workload, 4.0e9)
        return {"value": c / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc, "seconds": dt}
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from rnacode_b200 import capi
        self.torch, self.dist, self.capi, self.args = torch, dist, capi, args
        self.rank, self.world, self.local = dist_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.gloo = dist.new_group(backend="gloo")  # host-side gather of python objects (HSS records, maxima)
        else:
            self.gloo = None
        self.ctx = capi.Context(self.local)
        self.stream = torch.cuda.current_stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        self.prm = capi.make_params()
        self.blosum = np.array(BLOSUM62, dtype=np.int32)
        self.keep = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()

    # -- one workload: device-resident leg + e2e leg ---------------------------------------------------------------
    def measure(self, name, steps, warmup, host_samples, e2e_steps, n_override=0, sampler=None, pipelined=False):
        torch, capi = self.torch, self.capi
        blocks_np, n, seed, desc = build_workload(name, self.rank)
        if n_override:
            n = n_override
        cells = workload_cells(blocks_np, n)
        trees = [capi.Tree(*synth.synth_tree(seed, idx, rows.shape[0])) for rows, _, _, idx in blocks_np]
        seeds = [np.arange(1, n + 1, dtype=np.uint32) + np.uint32(7919 * i) for i in range(len(blocks_np))]
        blocks = []
        if host_samples:  # null alignments in pinned host memory (the e2e leg copies them every step)
            for rows, sf, sr, idx in blocks_np:
                N, cols = rows.shape
                t = torch.from_numpy(synth.synth_samples(seed, idx, n, N, cols)).pin_memory()
                r = torch.from_numpy(rows.copy()).pin_memory()
                self.keep += [t, r]
                blocks.append(capi.Block(r.numpy(), sf, sr, t.numpy()))
        else:  # null alignments drawn on the GPU (kernel d), the CLIs' default
            blocks = [capi.Block(rows, sf, sr, None, n_samples=n) for rows, sf, sr, _ in blocks_np]

        descs = capi.Batch.block_descs(blocks)                      # the caller's descriptor array and tree tables exist
        plan = None if host_samples else capi.Batch.evolve_plan(trees, seeds)  # before the step, like the rows themselves

        def make_batch():
            bt = self.ctx.batch(blocks, self.prm, self.blosum, descs)
            if not host_samples:
                bt.set_evolve_many(plan, capi.RC_RNG_MT19937)
            return bt

        # device-resident leg
        bt = make_batch()
        bt.upload()
        for _ in range(max(warmup, 3)):
            bt.run()
        self.barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = self.events()
        stage_ms = {"pack": 0.0, "sigma": 0.0, "dp": 0.0, "hss": 0.0}
        pack_kernel_ms, launches, dp_launches = 0.0, 0, 0
        self.barrier()
        ev0.record(self.stream)
        for _ in range(steps):
            bt.run()
            st = bt.stats()
            pack_kernel_ms += st["ms_pack_kernel"]
            for k in stage_ms:
                stage_ms[k] += st["ms_" + k]
            launches += st["launches"]
            dp_launches += st["dp_launches"]
        ev1.record(self.stream)
        self.barrier()
        total_ms = ev0.elapsed_time(ev1)
        if sampler:
            sampler.stop_flag.set()
            sampler.join(timeout=5)
        st = bt.stats()
        bt.download()
        n_hss = int(sum(len(bt.native_hss(i)) for i in range(len(blocks))))
        bt.close()

        # end-to-end leg through the C ABI with host buffers
        def e2e_step(ctx=None):
            b2 = make_batch() if ctx is None else make_batch_on(ctx)
            b2.upload()
            b2.run()
            b2.download()
            s2 = b2.stats()
            _ = b2.max_scores_all()
            b2.close()
            return s2

        def make_batch_on(ctx):
            bt = ctx.batch(blocks, self.prm, self.blosum, descs)
            if not host_samples:
                bt.set_evolve_many(plan, capi.RC_RNG_MT19937)
            return bt
        for _ in range(2):
            e2e_step()
        self.barrier()
        e0, e1 = self.events()
        e0.record(self.stream)
        for _ in range(e2e_steps):
            s2 = e2e_step()
        e1.record(self.stream)
        self.barrier()
        e2e_ms = e0.elapsed_time(e1)
        # the same steps the way a caller that streams batches runs them: two or three host threads, each with a context (and
        # stream) of its own, take the steps in turn, so that the host->device copy of one step overlaps the kernels of another.
        # Every step still uploads its own inputs from pinned host memory and reads its own results back inside the timed region.
        e2e_pipe_ms = None
        if pipelined:
            nthr = 2 if pipelined is True else int(pipelined)
            ctxs = [capi.Context(self.local) for _ in range(nthr)]
            errs = []

            def worker(k, n_steps):
                try:
                    for _ in range(n_steps):
                        e2e_step(ctxs[k])
                except Exception as e:  # surfaces after the join
                    errs.append(e)
            for nst in (nthr, e2e_steps):  # a warm-up round, then the timed one
                share = [nst // nthr + (1 if k < nst % nthr else 0) for k in range(nthr)]
                self.barrier()
                p0, p1 = self.events()
                p0.record(self.stream)
                th = [threading.Thread(target=worker, args=(k, share[k])) for k in range(nthr)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                torch.cuda.synchronize()  # the work of both contexts' streams is done before the closing event
                p1.record(self.stream)
                self.barrier()
                e2e_pipe_ms = p0.elapsed_time(p1)
            for c in ctxs:
                c.close()
            if errs:
                raise errs[0]
        return {"e2e_pipe_ms": e2e_pipe_ms,"name": name, "desc": desc, "blocks_np": blocks_np, "blocks": blocks, "trees": trees, "seeds": seeds, "n": n,
                "seed": seed, "cells": cells, "steps": steps, "total_ms": total_ms, "stage_ms": {k: v / steps for k, v in stage_ms.items()},
                "pack_kernel_ms": pack_kernel_ms / steps, "launches": launches, "dp_launches": dp_launches,
                "device_bytes": st["device_bytes"], "pack_chars": st["pack_chars"], "dense_fallbacks": st["dense_fallbacks"],
                "native_hss_total": n_hss, "e2e_ms": e2e_ms, "e2e_steps": e2e_steps, "h2d": int(s2["h2d_bytes"]),
                "d2h": int(s2["d2h_bytes"]), "host_samples": host_samples}

    def roofline(self, m, peak_nominal, peak_measured, sm_count, sm_max):
        dp_s = m["stage_ms"]["dp"] * 1e-3
        g = frameshift_fraction(m["blocks_np"][:24])
        achieved = m["cells"] * OPS_PER_CELL / dp_s
        blended = m["cells"] * (OPS_PER_CELL * (1 - g) + OPS_PER_CELL_FS * g) / dp_s
        return {"bound": "fp32_issue", "kernel": "k_dp*", "achieved": achieved / 1e12, "peak": peak_nominal / 1e12, "unit": "TFLOP/s",
                "frac": achieved / peak_nominal,
                "peak_source": "nominal %d SMs x 128 lanes x %.0f MHz (MEASURED_PEAKS.json holds no FP32-issue figure; see peak_measured)" % (
                    sm_count, sm_max),
                "peak_measured": peak_measured / 1e12,
                "peak_measured_source": "k_calib (rc_calibrate_issue): register-only loop of the DP cell's mix, 4 FADD : 1 FMNMX3, timed in this run",
                "frac_of_measured": achieved / peak_measured if peak_measured else None,
                "ops_per_cell": OPS_PER_CELL, "frameshift_fraction_g": g,
                "ops_per_cell_blended": OPS_PER_CELL * (1 - g) + OPS_PER_CELL_FS * g, "frac_blended": blended / peak_nominal,
                "kernel_ms_per_step": m["stage_ms"]["dp"], "stage_ms_per_step": m["stage_ms"]}

    # -- CLI end to end (host stages included) --------------------------------------------------------------------
    def cli_legs(self):
        exe = os.path.join(ROOT, "oracle", "_ref", "RNAcode_b200")
        maf = os.path.join(ROOT, "oracle", "_ref", "examples", "genomic.maf")
        if not (os.path.exists(exe) and os.path.exists(maf)):
            return {"unavailable": "oracle/_ref/RNAcode_b200 or its examples are not built"}
        out = {"host_cores": os.cpu_count(), "binary": "oracle/_ref/RNAcode_b200 (reference host stages + libRNAcode_cuda)"}
        env = dict(os.environ, RNACODE_CUDA_DEVICE=str(self.local), RNACODE_CUDA_VERBOSE="1")

        def run(cmd, reps):
            walls, log = [], ""
            for _ in range(reps):
                t0 = time.perf_counter()
                r = subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                walls.append(time.perf_counter() - t0)
                if r.returncode != 0:
                    return None, r.stderr[-300:]
                log = r.stderr
            stages = [ln.strip() for ln in log.splitlines() if ln.startswith("[RNAcode_b200] window")]
            return min(walls), stages[-1][:400] if stages else ""
        w, log = run([exe, "--gtf", "--best-only", "-n", "1000", maf], 2)
        if w is None:
            out["genomic_maf"] = {"error": log}
        else:
            nb = sum(1 for ln in open(maf) if ln.startswith("a"))
            out["genomic_maf"] = {"command": "RNAcode_b200 --gtf --best-only -n 1000 examples/genomic.maf", "wall_s": w, "blocks": nb,
                                  "blocks_per_s": nb / w, "stages": log,
                                  "note": "process start to exit, CUDA context creation (about 0.4 s) included; best of 2"}
        tmpdir = tempfile.mkdtemp(prefix="rc_cli_")
        try:
            nblk = int(os.environ.get("RC_BENCH_CLI_SHORT_BLOCKS", "10000"))
            p = os.path.join(tmpdir, "c3.maf")
            synth.to_maf([synth.synth_block(1, i, 10, 120) for i in range(nblk)], p)
            w, log = run([exe, "--tabular", "-n", "100", p], 1)
            if w is None:
                out["config3_maf"] = {"error": log}
            else:
                out["config3_maf"] = {"command": "RNAcode_b200 --tabular -n 100 <synthetic MAF %d blocks x 10 x 120>" % nblk, "wall_s": w,
                                      "blocks": nblk, "blocks_per_s": nblk / w, "stages": log}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
        return out

    # -- strong scaling through the sharder ---------------------------------------------------------------------
    def sharded(self, label, shapes, n, seed, stop_early, cutoff):
        """One fixed global block list, scored by all ranks together (timed) and by rank 0 alone (untimed check)."""
        capi = self.capi
        nb = len(shapes)
        cache = {}

        def block(i):
            if i not in cache:
                N, cols = shapes[i]
                rows = synth.synth_block(seed, i, N, cols)
                sf, sr = synth.synth_scores(seed, i, N)
                cache[i] = (rows, sf, sr, capi.Tree(*synth.synth_tree(seed, i, N)),
                            (np.arange(n, dtype=np.uint32) * np.uint32(2654435761) + np.uint32(40503 * i + 1)).astype(np.uint32))
            return cache[i]
        # the cost model needs every block's ungapped length, so every rank generates the whole list (a pure function of
        # (seed, index)); only the units a rank is dealt reach its GPU
        per_aln, cells_full = [], []
        for i, (N, cols) in enumerate(shapes):
            L = synth.ungapped_len(block(i)[0])
            per_aln.append(float(N - 1) * L * L)
            cells_full.append(synth.cells(N, L, 0))

        def make_scorer(ctx):
            def scorer(units):
                blks = [capi.Block(block(u.block)[0], block(u.block)[1], block(u.block)[2], None, n_samples=u.ns) for u in units]
                bt = ctx.batch(blks, self.prm, self.blosum)
                if all(u.ns > 0 for u in units):  # one call: the trees' tables are converted on several host threads
                    bt.set_evolve_many(capi.Batch.evolve_plan([block(u.block)[3] for u in units],
                                                              [block(u.block)[4][u.s0:u.s0 + u.ns] for u in units]),
                                       capi.RC_RNG_MT19937)
                else:
                    for k, u in enumerate(units):
                        if u.ns > 0:
                            bt.set_evolve(k, block(u.block)[3], block(u.block)[4][u.s0:u.s0 + u.ns], capi.RC_RNG_MT19937)
                bt.upload(); bt.run(); bt.download()
                allmax = bt.max_scores_all()
                res, pos = {}, 0
                for k, u in enumerate(units):
                    res[(u.block, u.s0)] = (bt.native_hss(k) if u.want_native else None, allmax[pos:pos + u.ns].copy())
                    pos += u.ns
                bt.close()
                return res
            return scorer
        scorer = make_scorer(self.ctx)
        run = lambda rank, world, group: shard.score_sharded(per_aln, n, rank, world, scorer, stop_early, cutoff, group)  # noqa: E731
        if self.world > 1:
            self.dist.barrier(group=self.gloo)
        run(self.rank, self.world, self.gloo)  # warm-up (allocations, first launches)
        self.barrier()
        e0, e1 = self.events()
        e0.record(self.stream)
        results, info = run(self.rank, self.world, self.gloo)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        vals = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(vals, op=self.dist.ReduceOp.MAX)
        ms = float(vals.item())
        out = None
        if self.rank == 0:
            # single-GPU scoring of the same list: the parity check, and the 1-GPU time of this very run
            old_world = self.world
            shard_dist_off = _NoDist()
            with shard_dist_off:
                run(0, 1, None)
                self.torch.cuda.synchronize()
                a0, a1 = self.events()
                a0.record(self.stream)
                single, _ = run(0, 1, None)
                a1.record(self.stream)
                self.torch.cuda.synchronize()
            t1 = a0.elapsed_time(a1)
            cells = 0.0
            for b, (hss, mx, status) in enumerate(results):
                cells += cells_full[b] * (1 + int(np.isfinite(mx).sum()))
            out = {"workload": label, "blocks": nb, "n_samples": n, "stop_early": bool(stop_early), "cells_scored": cells,
                   "ms": ms, "value": cells / (ms * 1e-3), "unit": UNIT, "n_gpus": old_world, "scaling": "strong",
                   "single_gpu_ms_same_run": t1, "speedup_vs_single_gpu": t1 / ms,
                   "shard_parity": shard.digest(results) == shard.digest(single), "digest": shard.digest(results)[:16],
                   "rounds": info["rounds"], "round2_blocks": info["round2_blocks"], "stopped_early": info["stopped_early"],
                   "native_hss_total": int(sum(len(r[0]) for r in results)),
                   "timed_region": "plan + rc_batch_create/set_evolve/upload/run/download per rank + host gather in input order"}
        if self.world > 1:
            self.dist.barrier(group=self.gloo)
        return out


class _NoDist:
    """Inside this context shard._gather does not talk to the other ranks (rank 0's single-GPU check)."""

    def __enter__(self):
        self._saved = shard._gather
        shard._gather = lambda local, group=None: [local]

    def __exit__(self, *a):
        shard._gather = self._saved


def sharded_lists():
    """The two fixed global block lists of the strong-scaling legs."""
    n4 = int(os.environ.get("RC_BENCH_C4_BLOCKS", "12"))
    n5 = int(os.environ.get("RC_BENCH_C5_BLOCKS", "4096"))
    rng = np.random.default_rng(3)
    c5 = [(100, int(round(float(np.exp(rng.uniform(np.log(60.0), np.log(2000.0))))))) for _ in range(n5)]
    return [("config 4: %d blocks x 50 species x 5000 cols, -n 1000" % n4, [(50, 5000)] * n4, 1000, 2, False, 1.0),
            ("config 5 sample: %d blocks x 100 species, cols log-uniform in [60, 2000] (numpy default_rng(3)), -n 100 -p 0.05 --stop-early" % n5,
             c5, 100, 3, True, 0.05)]


def ours(args):
    B = Bench(args)
    torch, dist = B.torch, B.dist
    rank, world, local = B.rank, B.world, B.local
    sampler = ClockSampler(local)
    m = B.measure(args.workload, args.steps, args.warmup, host_samples=not args.evolve, e2e_steps=args.steps,
                  n_override=args.samples, sampler=sampler, pipelined=3)
    if args.quick and rank == 0:  # kernel iteration: one compact line on stderr
        sys.stderr.write("[quick] %s: %.4g cells/s, %.3f ms/step, stages %s, e2e %.3f ms (pipelined %.3f), launches %d\n" % (
            args.workload, m["cells"] / (m["total_ms"] / m["steps"] * 1e-3), m["total_ms"] / m["steps"],
            {k: round(v, 3) for k, v in m["stage_ms"].items()}, m["e2e_ms"] / m["e2e_steps"],
            m["e2e_pipe_ms"] / m["e2e_steps"], m["launches"] // m["steps"]))

    # the same C-ABI sequence with the null alignments drawn on the GPU (kernel d, the CLIs' default): the host supplies the
    # native rows, the score tables, a flattened tree and one seed per sample
    capi = B.capi
    ev_blocks = [capi.Block(b.rows, b.scores_fwd, b.scores_rev, None, n_samples=m["n"]) for b in m["blocks"]]

    def evolve_step():
        b3 = B.ctx.batch(ev_blocks, B.prm, B.blosum)
        for i in range(len(ev_blocks)):
            b3.set_evolve(i, m["trees"][i], m["seeds"][i], capi.RC_RNG_MT19937)
        b3.upload(); b3.run(); b3.download()
        s3 = b3.stats()
        _ = [b3.max_scores(i) for i in range(len(ev_blocks))]
        b3.close()
        return s3
    for _ in range(2):
        evolve_step()
    B.barrier()
    ev_steps = max(1, min(args.steps, 5))
    v0, v1 = B.events()
    v0.record(B.stream)
    for _ in range(ev_steps):
        s3 = evolve_step()
    v1.record(B.stream)
    B.barrier()
    e2e_evolve_ms = v0.elapsed_time(v1)
    issue_measured = B.ctx.calibrate_issue()

    vals = torch.tensor([m["total_ms"], m["e2e_ms"], e2e_evolve_ms, m["e2e_pipe_ms"]], dtype=torch.float64, device="cuda")
    tot = torch.tensor([m["cells"], float(m["launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, e2e_evolve_ms, e2e_pipe_ms = [float(x) for x in vals.tolist()]
    cells_all, launches_all = [float(x) for x in tot.tolist()]

    line = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    clocks = sampler.summary()
    sm_max = peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    nominal_peak = sm_count * 128 * sm_max * 1e6  # lane-issues per second
    if rank == 0:
        ms_per_step = total_ms / args.steps
        nblocks = len(m["blocks"]) * world
        traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except Exception:
            pass
        roof = B.roofline(m, nominal_peak, issue_measured, sm_count, sm_max)
        roof["traffic"] = traffic["dram_bytes_per_launch"] if traffic and not args.samples else None
        roof["traffic_source"] = traffic["source"] if traffic and not args.samples else None
        line = {
            "metric": METRIC, "value": cells_all / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, args.samples),
            "device_bytes_per_step": int(m["device_bytes"]),
            "blocks_per_s": nblocks / (ms_per_step * 1e-3),
            "clocks": clocks,
            "e2e": {"value": cells_all / (e2e_pipe_ms / m["e2e_steps"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": e2e_pipe_ms / m["e2e_steps"], "steps": m["e2e_steps"],
                    "blocks_per_s": nblocks / (e2e_pipe_ms / m["e2e_steps"] * 1e-3),
                    "how": "C ABI from pinned host buffers, every step: rc_batch_create + upload (H2D) + run + download (D2H) + "
                           "destroy; three host threads with a context and stream each take the steps in turn, so the copies "
                           "and the host work of one step overlap the kernels of another (what a caller streaming batches does)",
                    "serial": {"value": cells_all / (e2e_ms / m["e2e_steps"] * 1e-3), "ms_per_step": e2e_ms / m["e2e_steps"],
                               "how": "the same steps one after the other on one context: copies and kernels never overlap"}},
            "e2e_gpu_evolve": {"value": cells_all / (e2e_evolve_ms / ev_steps * 1e-3), "unit": UNIT,
                               "h2d_bytes_per_step": int(s3["h2d_bytes"]), "d2h_bytes_per_step": int(s3["d2h_bytes"]),
                               "ms_per_step": e2e_evolve_ms / ev_steps,
                               "note": "same C-ABI sequence with the null alignments simulated on the GPU (exact MT19937 mode)"},
            "gpu_launches": int(launches_all),
            "roofline": roof,
            # subsystem (a): k_pack against the measured HBM copy bandwidth; algorithmic bytes per character in DESIGN.md section 4
            "roofline_pack": {"bound": "hbm", "kernel": "k_pack", "unit": "GB/s",
                              "achieved": PACK_BYTES_PER_CHAR * m["pack_chars"] / (m["pack_kernel_ms"] * 1e-3) / 1e9 if m["pack_kernel_ms"] else None,
                              "peak": peaks.get("hbm_gbs"),
                              "frac": (PACK_BYTES_PER_CHAR * m["pack_chars"] / (m["pack_kernel_ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"])
                              if m["pack_kernel_ms"] and peaks.get("hbm_gbs") else None,
                              "bytes_per_char": PACK_BYTES_PER_CHAR, "kernel_ms_per_step": m["pack_kernel_ms"]},
            "dense_fallbacks": int(m["dense_fallbacks"]),
            "native_hss_total": int(m["native_hss_total"]),
        }
    del m["blocks"], m["blocks_np"]
    B.keep.clear()

    # the other BASELINE configs (N = 1 only: they describe one GPU)
    if world == 1 and not args.no_side:
        side = {}
        for name in SIDE_WORKLOADS:
            try:
                s = B.measure(name, steps=5, warmup=3, host_samples=False, e2e_steps=6, pipelined=3)
                r = B.roofline(s, nominal_peak, issue_measured, sm_count, sm_max)
                ms = s["total_ms"] / s["steps"]
                side[name] = {"workload": s["desc"], "blocks": len(s["blocks"]), "n_samples": s["n"], "cells_per_step": s["cells"],
                              "value": s["cells"] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": s["steps"],
                              "blocks_per_s": len(s["blocks"]) / (ms * 1e-3), "stage_ms_per_step": s["stage_ms"],
                              "roofline_frac": r["frac"], "roofline_frac_blended": r["frac_blended"],
                              "frameshift_fraction_g": r["frameshift_fraction_g"], "gpu_launches": int(s["launches"]),
                              "e2e": {"value": s["cells"] / (s["e2e_pipe_ms"] / s["e2e_steps"] * 1e-3), "unit": UNIT,
                                      "ms_per_step": s["e2e_pipe_ms"] / s["e2e_steps"], "h2d_bytes_per_step": s["h2d"],
                                      "d2h_bytes_per_step": s["d2h"],
                                      "note": "C ABI from host buffers: native rows, score tables, trees, seeds; null alignments drawn "
                                              "on the GPU; three host threads / contexts take the steps in turn (the headline's e2e uses two)",
                                      "serial": {"value": s["cells"] / (s["e2e_ms"] / s["e2e_steps"] * 1e-3),
                                                 "ms_per_step": s["e2e_ms"] / s["e2e_steps"]}},
                              "null_alignments": "drawn on the GPU inside the step (k_evolve, exact MT19937 mode)"}
                del s
            except Exception as e:  # a side workload must not take the headline down
                side[name] = {"error": repr(e)[:300]}
        line["workloads"] = side
        if not args.no_cli:
            line["blocks_per_s_e2e_cli"] = B.cli_legs()

    # strong scaling through the sharder (every N, so that the series N = 1, 2, 4, 8 holds the same lists)
    if not args.no_sharded:
        sh = []
        for label, shapes, n, seed, stop, cutoff in sharded_lists():
            try:
                r = B.sharded(label, shapes, n, seed, stop, cutoff)
            except Exception as e:
                r = {"workload": label, "error": repr(e)[:300]}
            sh.append(r)
        if rank == 0:
            line["sharded"] = sh
            line["shard_parity"] = all(bool(r.get("shard_parity")) for r in sh)
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line))
    B.close()
    return 0


PACK_BYTES_PER_CHAR = 2.0  # k_pack: 16 B sample read + 16 B class bytes written per 16 characters (the native row is L2-resident)

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
    ap.add_argument("--no-side", action="store_true", help="skip the other BASELINE configs (workloads) and the CLI leg")
    ap.add_argument("--no-cli", action="store_true", help="skip the CLI end-to-end leg")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling legs through the sharder")
    ap.add_argument("--evolve", action="store_true", help="headline legs with the null alignments drawn on the GPU instead of host buffers")
    ap.add_argument("--quick", action="store_true", help="kernel iteration: headline legs only (= --no-cpu --no-side --no-sharded)")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu = args.no_side = args.no_sharded = True
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
