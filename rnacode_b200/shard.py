"""Sharding of independent alignment blocks over the GPUs of one box.

Blocks carry nothing from one to the next (src/RNAcode.c:115-221) except the running hit counter of the
printer, so the data path needs no collective: every rank scores its own units, and the per-block results
are gathered on the host and re-serialised in input order.  torch.distributed is used for that gather only.

This module is the Python mirror of the sharder inside the batched CLI (integration/rnacode_pipeline.c,
gpu_batch / process_window): the same unit of work (a block's native alignment plus a range of its null
alignments), the same cost model, the same cut of oversize blocks along their null alignments, the same
heaviest-first assignment and the same two-round --stop-early sampling.  bench.py --gpus N drives it with
libRNAcode_cuda as the scorer; tests/test_shard_gloo.py drives it with the CPU oracle.
"""
import hashlib
import heapq
from collections import namedtuple

import numpy as np

from . import synth

# want_native: this unit also reports the block's native HSS list (the first part of a cut block)
Unit = namedtuple("Unit", "block s0 ns want_native cost")


def block_cost(N, L, n_samples):
    """Work of one block in DP cells: (n+1) * 2 * (N-1) * P(L)."""
    return synth.cells(N, L, n_samples)


def plan(costs, world_size):
    """Longest-processing-time-first assignment.  Returns a list (per rank) of block indices, each list in
    input order.  Deterministic: ties broken by block index, then by rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + float(costs[i]), r))
    return [sorted(s) for s in shards]


def loads(costs, shards):
    return [sum(costs[i] for i in s) for s in shards]


def plan_units(per_aln, blocks, s0, ns, world_size, want_native=True):
    """Units of GPU work for the null alignments [s0, s0+ns) (and, if want_native, the native alignment) of the listed
    blocks, dealt out over world_size devices -- gpu_batch() of integration/rnacode_pipeline.c:
    per_aln[i] = (N-1) * L * L is the cost of one alignment of block i; a block whose cost alone exceeds half a device's
    fair share is cut along its null alignments into world_size parts (the native alignment goes with the first part);
    units are assigned heaviest first, each to the device with the least work so far.
    Returns (per-rank lists of Unit, per-rank loads)."""
    G = world_size
    total = sum(per_aln[i] * (ns + 1) for i in blocks)
    units = []
    for i in blocks:
        parts = G if (G > 1 and ns >= 2 * G and per_aln[i] * (ns + 1) > total / (2.0 * G)) else 1
        for p in range(parts):
            a, e = ns * p // parts, ns * (p + 1) // parts
            units.append(Unit(i, s0 + a, e - a, bool(want_native and p == 0), float(per_aln[i]) * (e - a + 1)))
    order = sorted(range(len(units)), key=lambda k: (-units[k].cost, k))  # stable: equal costs keep input order
    shards = [[] for _ in range(G)]
    load = [0.0] * G
    heap = [(0.0, r) for r in range(G)]  # (work so far, device): the least loaded device, lowest index on ties
    for k in order:
        _, d = heapq.heappop(heap)
        shards[d].append(units[k])
        load[d] += units[k].cost
        heapq.heappush(heap, (load[d], d))
    return shards, load


def _gather(local, group=None):
    """all_gather_object of a python object (host-side gather; no collective on the data path)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    return parts


def gather_in_order(local, n_blocks, group=None):
    """local: {block index: result}.  Returns on every rank the list of results in input order
    (host-side gather of HSS records; all_gather_object so that any rank may print)."""
    merged = {}
    for p in _gather(local, group):
        for k, v in p.items():
            if k in merged:
                raise ValueError("block %d scored by two ranks" % k)
            merged[k] = v
    missing = [i for i in range(n_blocks) if i not in merged]
    if missing:
        raise ValueError("blocks not scored by any rank: %s" % missing[:8])
    return [merged[i] for i in range(n_blocks)]


def score_sharded(per_aln, n, rank, world_size, scorer, stop_early=False, cutoff=1.0, group=None, first_round=32):
    """The scoring part of process_window() (integration/rnacode_pipeline.c) over world_size ranks.

    scorer(units) scores this rank's units and returns {(block, s0): (native HSS list or None, maxima of the unit's null
    alignments as float64 array)}.  Without stop_early every block gets its n null alignments in one round; with it, a first
    round of `first_round` null alignments decides most non-coding blocks (more than int(cutoff*n) of them beat the best
    native score, src/score.c:992, :1036-1042 -- the count is monotone in the sample index, so the verdicts are the
    reference's) and only the others get the remaining n - first_round.
    Returns on every rank, in input order: [(native HSS list, maxima float64[n] (unsampled entries NaN), status)], status 1 =
    all n null alignments scored, -1 = stopped early; plus a dict of planning facts (units, loads per round)."""
    nb = len(per_aln)
    n1 = first_round if (stop_early and n > first_round) else n
    stop_cut = int(cutoff * n)
    hss = [None] * nb
    M = np.full((nb, n), np.nan)  # maxima of the null alignments, NaN = not sampled
    maxima = [M[b] for b in range(nb)]  # row views
    info = {"rounds": []}

    def one_round(blocks, s0, ns, want_native):
        shards, load = plan_units(per_aln, blocks, s0, ns, world_size, want_native)
        local = scorer(shards[rank]) if shards[rank] else {}
        for part in _gather(local, group):
            for (b, u0), (h, mx) in part.items():
                if h is not None:
                    if hss[b] is not None:
                        raise ValueError("native alignment of block %d scored twice" % b)
                    hss[b] = h
                if np.isfinite(maxima[b][u0:u0 + len(mx)]).any():
                    raise ValueError("null alignments of block %d scored twice" % b)
                maxima[b][u0:u0 + len(mx)] = mx
        info["rounds"].append({"blocks": len(blocks), "units": sum(len(s) for s in shards), "samples": ns,
                               "load_max_over_mean": (max(load) * world_size / sum(load)) if sum(load) > 0 else 1.0})

    one_round(list(range(nb)), 0, n1, True)
    missing = [b for b in range(nb) if hss[b] is None]
    if missing:
        raise ValueError("blocks not scored by any rank: %s" % missing[:8])
    # results[0].score after the sort (src/RNAcode.c:176-178), -1 when the list is empty (src/score.c:1129-1134)
    best = np.array([max([h[4] for h in hss[b]], default=-1.0) for b in range(nb)], dtype=np.float32)

    def verdicts(upto):
        """src/score.c:1036-1042 over the first `upto` null alignments: -1 as soon as more than stop_cut of them beat the
        native score (the running count is monotone, so "at some point" == "at the end of the prefix")."""
        better = (M[:, :upto].astype(np.float32) > best[:, None]).sum(axis=1)
        return np.where(stop_early & (better > stop_cut), -1, 1)
    status = verdicts(n1)
    todo = [b for b in range(nb) if status[b] == 1] if n1 < n else []
    if todo:
        one_round(todo, n1, n - n1, False)
        status = np.where(status == 1, verdicts(n), status)
    full = status == 1
    if not np.isfinite(M[full]).all():
        raise ValueError("null alignments missing for some fully sampled block")
    info["round2_blocks"] = len(todo)
    info["stopped_early"] = int((status < 0).sum())
    return [(hss[b], M[b], int(status[b])) for b in range(nb)], info


def digest(results):
    """SHA-256 over everything the reporting stage consumes: per block the native HSS records (strand, frame, sites, float32
    score bits), the status, and the float32 bits of the sample maxima that the verdict depends on (all n when the block was
    sampled in full; for a block stopped early only the decisive prefix is defined by the reference, so only the status)."""
    h = hashlib.sha256()
    for hss, mx, status in results:
        h.update(np.int32(len(hss)).tobytes())
        for s, f, a, e, sc in hss:
            h.update(("%s%d:%d:%d:" % (s, f, a, e)).encode())
            h.update(np.float32(sc).tobytes())
        h.update(np.int32(status).tobytes())
        if status == 1:
            h.update(np.asarray(mx, dtype=np.float32).tobytes())
    return h.hexdigest()
