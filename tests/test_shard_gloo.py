"""Multi-process (world_size 2, gloo, CPU) test of the N>1 host logic (rnacode_b200/shard.py, the mirror of the sharder in
integration/rnacode_pipeline.c): cost-weighted units of (block, range of null alignments), oversize blocks cut along their
null alignments, independent scoring per rank, host-side gather in input order, two-round --stop-early sampling.  The
per-rank scorer here is the CPU oracle (this is a test); on the GPU box bench.py --gpus N drives the same code with
libRNAcode_cuda and compares the same digest."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_SAMPLES = 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _blocks():
    from rnacode_b200 import synth
    # one block far heavier than the rest: it is cut along its null alignments when there are two ranks
    shapes = [(6, 90), (10, 600), (4, 45), (8, 150), (5, 60), (12, 120), (3, 30)]
    out = []
    for i, (N, cols) in enumerate(shapes):
        rows = synth.synth_block(31, i, N, cols, gap_rate=0.02)
        sf, sr = synth.synth_scores(31, i, N)
        smp = synth.synth_samples(31, i, N_SAMPLES, N, cols)
        out.append((rows, sf, sr, smp))
    return out


def _oracle_scorer(orc, blocks):
    prm = orc.params()

    def scorer(units):
        res = {}
        for u in units:
            rows, sf, sr, smp = blocks[u.block]
            hss = orc.score_aln(rows, sf, sr, prm) if u.want_native else None
            res[(u.block, u.s0)] = (hss, orc.sample_maxima(rows, smp[u.s0:u.s0 + u.ns], sf, sr, prm))
        return res
    return scorer


def _per_aln(blocks):
    from rnacode_b200 import synth
    return [float(b[0].shape[0] - 1) * synth.ungapped_len(b[0]) ** 2 for b in blocks]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rnacode_b200 import shard
    from tests import oracle_py as op
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    blocks = _blocks()
    orc = op.load()
    out = {}
    for stop in (False, True):
        res, info = shard.score_sharded(_per_aln(blocks), N_SAMPLES, rank, world, _oracle_scorer(orc, blocks), stop_early=stop,
                                        cutoff=0.2, first_round=16)
        out[stop] = (shard.digest(res), info, [(r[0], r[2]) for r in res])
    if rank == 0:
        q.put(out)
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


def test_plan_units_cuts_oversize_blocks():
    """gpu_batch() of integration/rnacode_pipeline.c: a block heavier than half a device's fair share is cut into one part
    per device along its null alignments; the native alignment goes with the first part; every sample is covered once."""
    from rnacode_b200 import shard
    per_aln = [1000.0, 10.0, 10.0, 12.0, 9.0]
    for world in (1, 2, 4, 8):
        shards, load = shard.plan_units(per_aln, list(range(5)), 0, 100, world)
        units = [u for s in shards for u in s]
        for b in range(5):
            mine = sorted((u for u in units if u.block == b), key=lambda u: u.s0)
            assert sum(u.want_native for u in mine) == 1 and mine[0].want_native
            pos = 0
            for u in mine:
                assert u.s0 == pos
                pos += u.ns
            assert pos == 100
            assert len(mine) == (world if (b == 0 and world > 1) else 1)
        assert abs(sum(load) - sum(u.cost for u in units)) < 1e-6
        if world > 1:
            assert max(load) / (sum(load) / world) < 1.15
    # second round of --stop-early: only the listed blocks, samples from s0 on, no native alignment
    shards, _ = shard.plan_units(per_aln, [1, 3], 32, 68, 2, want_native=False)
    units = [u for s in shards for u in s]
    assert sorted(set(u.block for u in units)) == [1, 3] and not any(u.want_native for u in units)
    for b in (1, 3):  # with only two blocks left each outweighs half a device's share: two parts each, covering [32, 100)
        mine = sorted((u for u in units if u.block == b), key=lambda u: u.s0)
        assert [(u.s0, u.ns) for u in mine] == [(32, 34), (66, 34)]


def test_two_ranks_match_single_process():
    from rnacode_b200 import shard
    from tests import oracle_py as op
    op.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = op.load()
    blocks = _blocks()
    for stop in (False, True):
        single, info1 = shard.score_sharded(_per_aln(blocks), N_SAMPLES, 0, 1, _oracle_scorer(orc, blocks), stop_early=stop,
                                            cutoff=0.2, first_round=16)
        dig, info2, brief = out[stop]
        assert dig == shard.digest(single), stop
        assert brief == [(r[0], r[2]) for r in single]
        assert info2["rounds"][0]["units"] == len(blocks) + 1  # the 10 x 600 block went out in two parts
        assert info1["rounds"][0]["units"] == len(blocks)
        # plain oracle calls, block by block
        for i, (rows, sf, sr, smp) in enumerate(blocks):
            assert single[i][0] == orc.score_aln(rows, sf, sr, orc.params())
            if single[i][2] == 1:
                exp = orc.sample_maxima(rows, smp, sf, sr, orc.params())
                assert np.array_equal(single[i][1].astype(np.float32), exp.astype(np.float32))
    stopped = [r[2] for r in shard.score_sharded(_per_aln(blocks), N_SAMPLES, 0, 1, _oracle_scorer(orc, blocks), stop_early=True,
                                                 cutoff=0.2, first_round=16)[0]]
    assert -1 in stopped  # the synthetic blocks are not coding: most stop after the first round
