"""Multi-process (world_size 2, gloo, CPU) test of the N>1 host logic: cost-weighted sharding of blocks,
independent scoring per rank, host-side gather in input order.  The per-rank scorer here is the CPU oracle
(this is a test); on the GPU box bench.py / the tests run the same plan with libRNAcode_cuda."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _blocks():
    from rnacode_b200 import synth
    shapes = [(6, 90, 3), (10, 240, 2), (4, 45, 4), (8, 150, 2), (5, 60, 3), (12, 120, 2), (3, 30, 5)]
    out = []
    for i, (N, cols, n) in enumerate(shapes):
        rows = synth.synth_block(31, i, N, cols, gap_rate=0.02)
        sf, sr = synth.synth_scores(31, i, N)
        smp = synth.synth_samples(31, i, n, N, cols)
        out.append((rows, sf, sr, smp))
    return out


def _score(orc, blk):
    rows, sf, sr, smp = blk
    prm = orc.params()
    return orc.score_aln(rows, sf, sr, prm), [float(x) for x in orc.sample_maxima(rows, smp, sf, sr, prm)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rnacode_b200 import shard, synth
    from tests import oracle_py as op
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    blocks = _blocks()
    costs = [shard.block_cost(b[0].shape[0], synth.ungapped_len(b[0]), b[3].shape[0]) for b in blocks]
    shards = shard.plan(costs, world)
    orc = op.load()
    local = {i: _score(orc, blocks[i]) for i in shards[rank]}
    allres = shard.gather_in_order(local, len(blocks))
    if rank == 0:
        q.put((shards, allres))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_is_balanced_and_complete():
    from rnacode_b200 import shard
    costs = [100, 1, 1, 50, 49, 1, 98, 3]
    for w in (1, 2, 3, 8):
        sh = shard.plan(costs, w)
        assert sorted(i for s in sh for i in s) == list(range(len(costs)))
        ld = shard.loads(costs, sh)
        assert max(ld) <= sum(costs) / w + max(costs)
    assert shard.plan(costs, 2) == shard.plan(costs, 2)


def test_two_ranks_match_single_process():
    from tests import oracle_py as op
    op.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    shards, allres = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(len(s) > 0 for s in shards)
    orc = op.load()
    blocks = _blocks()
    for i, blk in enumerate(blocks):
        hss, mx = _score(orc, blk)
        assert allres[i][0] == hss
        assert np.array_equal(np.array(allres[i][1], dtype=np.float32), np.array(mx, dtype=np.float32))
