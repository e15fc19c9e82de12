"""Sharding of independent alignment blocks over the GPUs of one box.

Blocks carry nothing from one to the next (src/RNAcode.c:115-221) except the running hit counter of the
printer, so the data path needs no collective: every rank scores its own blocks, and the per-block results
are gathered on the host and re-serialised in input order.  torch.distributed is used for that gather only.
"""
import heapq

from . import synth


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


def gather_in_order(local, n_blocks, group=None):
    """local: {block index: result}.  Returns on every rank the list of results in input order
    (host-side gather of HSS records; all_gather_object so that any rank may print)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        parts = [local]
    else:
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, local, group=group)
    merged = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise ValueError("block %d scored by two ranks" % k)
            merged[k] = v
    missing = [i for i in range(n_blocks) if i not in merged]
    if missing:
        raise ValueError("blocks not scored by any rank: %s" % missing[:8])
    return [merged[i] for i in range(n_blocks)]
