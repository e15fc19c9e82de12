"""Wall-clock of each C-ABI call of one batch (host + device, synchronised) for a bench workload:
python tools/time_batch_host.py short"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from rnacode_b200 import capi, synth

w = sys.argv[1] if len(sys.argv) > 1 else "short"
blocks_np, n, seed, desc = bench.build_workload(w, 0)
blocks = []
for rows, sf, sr, idx in blocks_np:
    N, cols = rows.shape
    t_ = torch.from_numpy(synth.synth_samples(seed, idx, n, N, cols)).pin_memory()
    r_ = torch.from_numpy(rows.copy()).pin_memory()
    keep = globals().setdefault("keep", [])
    keep += [t_, r_]
    blocks.append(capi.Block(r_.numpy(), sf, sr, t_.numpy()))
ctx = capi.Context(0)
prm = capi.make_params()
blosum = np.array(bench.BLOSUM62, dtype=np.int32)
for rep in range(4):
    t = [time.perf_counter()]
    b = ctx.batch(blocks, prm, blosum); t.append(time.perf_counter())
    b.upload(); torch.cuda.synchronize(); t.append(time.perf_counter())
    b.run(); torch.cuda.synchronize(); t.append(time.perf_counter())
    b.download(); t.append(time.perf_counter())
    _ = [b.max_scores(i) for i in range(len(blocks))]; t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    names = ["create", "upload", "run", "download", "max_scores", "close"]
    print(desc[:40], " ".join("%s %.2f" % (nm, (t[i + 1] - t[i]) * 1e3) for i, nm in enumerate(names)), "ms  total %.2f" % ((t[-1] - t[0]) * 1e3))

# the same with the null alignments drawn on the GPU (what the CLIs and bench.py's side workloads do)
trees = [capi.Tree(*synth.synth_tree(seed, idx, rows.shape[0])) for rows, _, _, idx in blocks_np]
seeds = [np.arange(1, n + 1, dtype=np.uint32) + np.uint32(7919 * i) for i in range(len(blocks))]
ev_blocks = [capi.Block(b.rows, b.scores_fwd, b.scores_rev, None, n_samples=n) for b in blocks]
descs = capi.Batch.block_descs(ev_blocks)
plan = capi.Batch.evolve_plan(trees, seeds)
for rep in range(4):
    t = [time.perf_counter()]
    b = ctx.batch(ev_blocks, prm, blosum, descs); t.append(time.perf_counter())
    b.set_evolve_many(plan, capi.RC_RNG_MT19937); t.append(time.perf_counter())
    b.upload(); torch.cuda.synchronize(); t.append(time.perf_counter())
    b.run(); torch.cuda.synchronize(); t.append(time.perf_counter())
    b.download(); t.append(time.perf_counter())
    _ = b.max_scores_all(); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    names = ["create", "set_evolve_many", "upload", "run", "download", "max_scores_all", "close"]
    print("evolve:", " ".join("%s %.2f" % (nm, (t[i + 1] - t[i]) * 1e3) for i, nm in enumerate(names)), "ms  total %.2f" % ((t[-1] - t[0]) * 1e3))
