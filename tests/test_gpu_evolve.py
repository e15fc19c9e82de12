"""Kernel (d): null alignments simulated on the GPU.

Exact mode (MT19937): with the seeds the reference used (oracle/ref_wrap.c) the GPU must redraw the reference's
simulated alignments byte for byte (golden `samples`, produced by seq-gen inside the unmodified reference) and
hence the same per-sample maxima.  GPU-RNG mode (Philox): same distribution -- two-sample KS test on the per-sample
maxima against the exact mode -- and reproducible for fixed seeds."""
import numpy as np
import pytest

from tests import oracle_py as op

pytestmark = pytest.mark.gpu


def _tree(capi, blk):
    nodes = blk["evolve"]["nodes"]
    cum = np.array([n["cum"] for n in nodes], dtype=np.float64)
    cum[0, :4] = blk["evolve"]["addFreq"]  # node 0 carries the cumulative root frequencies
    return capi.Tree([n["parent"] for n in nodes], [n["row"] for n in nodes], cum)


@pytest.mark.parametrize("name", ["coding_aln", "noncoding_aln", "coding_maf", "genomic_maf", "genomic_pre_maf",
                                  "synth_gappy", "synth_gapfree"])
def test_exact_mode_reproduces_seqgen(rc_ctx, name):
    from rnacode_b200 import capi
    doc = op.golden(name)
    prm = capi.make_params(**op.golden_params(doc))
    blocks, blks = [], []
    for blk in doc["blocks"]:
        if blk.get("skipped") or not blk.get("samples"):
            continue
        rows, sf, sr, smp = op.block_arrays(doc, blk)
        blocks.append(capi.Block(rows, sf, sr, None, n_samples=len(smp)))
        blks.append((blk, smp))
    bt = rc_ctx.batch(blocks, prm, doc["blosum"])
    for i, (blk, smp) in enumerate(blks):
        bt.set_evolve(i, _tree(capi, blk), blk["seeds"][:len(smp)], capi.RC_RNG_MT19937)
    bt.upload(); bt.run(); bt.download()
    for i, (blk, smp) in enumerate(blks):
        for s in range(len(smp)):
            assert np.array_equal(bt.sample_rows(i, s), smp[s]), (name, blk["index"], s)
        got = bt.max_scores(i).astype(np.float32)
        assert np.array_equal(got, np.array(blk["maxScores"][:len(smp)], dtype=np.float32)), (name, blk["index"])
        assert bt.native_hss(i) == op.expected_hss(blk)
    bt.close()


def test_genomic_maf_full_n1000_maxima(rc_ctx):
    """BASELINE config 2 as quoted: examples/genomic.maf at -n 1000.  The unmodified reference (ref_probe, 14 core-minutes,
    tests/golden/make_golden.py:long_cases) gives, per block, the native HSS list and the best score of each of its 1000
    null alignments (src/score.c:1004-1044); the library redraws those alignments from the same seeds and tree (kernel d)
    and must reproduce all 10 x 1000 maxima and every HSS bit for bit -- the 10 x 4806 block through k_dp_reg, the short
    ones through the sample-major kernels."""
    from rnacode_b200 import capi
    doc = op.golden("genomic_maf_n1000")
    prm = capi.make_params(**op.golden_params(doc))
    blocks, blks = [], []
    for blk in doc["blocks"]:
        if blk.get("skipped"):
            continue
        rows, sf, sr, _ = op.block_arrays(doc, blk)
        assert len(blk["maxScores"]) == len(blk["seeds"]) == 1000
        blocks.append(capi.Block(rows, sf, sr, None, n_samples=1000))
        blks.append(blk)
    assert len(blocks) == 10 and max(b.cols for b in blocks) == 4806
    bt = rc_ctx.batch(blocks, prm, doc["blosum"])
    for i, blk in enumerate(blks):
        bt.set_evolve(i, _tree(capi, blk), blk["seeds"], capi.RC_RNG_MT19937)
    bt.upload(); bt.run(); bt.download()
    for i, blk in enumerate(blks):
        assert bt.native_hss(i) == op.expected_hss(blk), blk["index"]
        got = bt.max_scores(i).astype(np.float32)
        assert np.array_equal(got, np.array(blk["maxScores"], dtype=np.float32)), blk["index"]
    assert bt.stats()["dense_fallbacks"] == 0
    bt.close()


def test_oracle_evolve_matches_gpu_on_long_rows(rc_ctx, oracle):
    """Rows longer than one MT19937 batch (624 draws) and not a multiple of 32: batch-boundary handling."""
    import ctypes as C
    from rnacode_b200 import capi, synth
    doc = op.golden("coding_aln")
    blk = doc["blocks"][0]
    tree = _tree(capi, blk)
    N = blk["N"]
    for cols in (31, 625, 1301):
        rows = synth.synth_block(3, cols, N, cols, gap_rate=0.01)
        sf, sr = synth.synth_scores(3, cols, N)
        seeds = np.array([17, 4000000000, 123456789], dtype=np.uint32)
        b = capi.Block(rows, sf, sr, None, n_samples=len(seeds))
        bt = rc_ctx.batch([b], capi.make_params(), doc["blosum"])
        bt.set_evolve(0, tree, seeds)
        bt.upload(); bt.run(); bt.download()
        oracle.lib.orc_evolve.argtypes = [C.c_ulong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p]
        af = np.ascontiguousarray(tree.cum[0, :4])
        smp = []
        for si, seed in enumerate(seeds):
            out = np.zeros((N, cols), dtype=np.uint8)
            oracle.lib.orc_evolve(int(seed), len(tree.parent), tree.parent.ctypes.data, tree.row.ctypes.data,
                                  tree.cum.ctypes.data, af.ctypes.data, N, cols, out.ctypes.data)
            assert np.array_equal(bt.sample_rows(0, si), out), (cols, si)
            smp.append(out)
        exp = oracle.sample_maxima(rows, np.stack(smp), sf, sr, oracle.params()).astype(np.float32)
        assert np.array_equal(bt.max_scores(0).astype(np.float32), exp)
        bt.close()


def test_philox_mode_same_distribution(rc_ctx):
    from scipy import stats
    from rnacode_b200 import capi
    doc = op.golden("coding_aln")
    blk = doc["blocks"][0]
    rows, sf, sr, _ = op.block_arrays(doc, blk)
    prm = capi.make_params(**op.golden_params(doc))
    n = 600
    tree = _tree(capi, blk)
    res = {}
    for rng, seeds in ((capi.RC_RNG_MT19937, np.arange(1, n + 1)), (capi.RC_RNG_PHILOX, np.arange(10001, 10001 + n)),
                       ("again", np.arange(10001, 10001 + n))):
        mode = capi.RC_RNG_PHILOX if rng == "again" else rng
        b = capi.Block(rows, sf, sr, None, n_samples=n)
        bt = rc_ctx.batch([b], prm, doc["blosum"])
        bt.set_evolve(0, tree, seeds.astype(np.uint32), mode)
        bt.upload(); bt.run(); bt.download()
        res[rng] = bt.max_scores(0)
        if rng == capi.RC_RNG_PHILOX:
            smp = np.stack([bt.sample_rows(0, s) for s in range(50)])
            # simulated rows are over ACGT and the reference row follows the stationary frequencies
            assert set(np.unique(smp)) <= set(b"ACGT")
            freq = np.array([(smp[:, 0, :] == c).mean() for c in b"ACGT"])
            af = np.diff(np.concatenate([[0.0], blk["evolve"]["addFreq"]]))
            assert np.all(np.abs(freq - af) < 0.03), (freq, af)
        bt.close()
    assert np.array_equal(res[capi.RC_RNG_PHILOX], res["again"])
    ks = stats.ks_2samp(res[capi.RC_RNG_MT19937], res[capi.RC_RNG_PHILOX])
    assert ks.pvalue > 1e-3, ks


def test_set_evolve_rejects_bad_trees(rc_ctx):
    from rnacode_b200 import capi
    doc = op.golden("coding_aln")
    blk = doc["blocks"][0]
    rows, sf, sr, _ = op.block_arrays(doc, blk)
    tree = _tree(capi, blk)
    b = capi.Block(rows, sf, sr, None, n_samples=2)
    bt = rc_ctx.batch([b], capi.make_params(), doc["blosum"])
    bad = capi.Tree(tree.parent[::-1].copy(), tree.row, tree.cum)
    with pytest.raises(capi.RcError):
        bt.set_evolve(0, bad, [1, 2])
    with pytest.raises(capi.RcError):
        bt.upload()  # samples missing and no tree
    bt.close()


def test_bulk_calls_match_per_block_calls(rc_ctx):
    """rc_batch_set_evolve_many / rc_batch_max_scores_all (one call for a window of blocks) give what the per-block calls give."""
    from rnacode_b200 import capi
    doc = op.golden("genomic_pre_maf")
    prm = capi.make_params(**op.golden_params(doc))
    blks = [b for b in doc["blocks"] if not b.get("skipped")][:12]
    blocks = [capi.Block(*op.block_arrays(doc, b)[:3], None, n_samples=24 + (i % 5)) for i, b in enumerate(blks)]
    trees = [_tree(capi, b) for b in blks]
    seeds = [np.arange(100 * i + 1, 100 * i + 1 + blocks[i].n_samples, dtype=np.uint32) for i in range(len(blks))]
    one = rc_ctx.batch(blocks, prm, doc["blosum"])
    for i in range(len(blocks)):
        one.set_evolve(i, trees[i], seeds[i], capi.RC_RNG_MT19937)
    one.upload(); one.run(); one.download()
    many = rc_ctx.batch(blocks, prm, doc["blosum"], capi.Batch.block_descs(blocks))
    many.set_evolve_many(capi.Batch.evolve_plan(trees, seeds), capi.RC_RNG_MT19937)
    many.upload(); many.run(); many.download()
    assert np.array_equal(many.max_scores_all(), np.concatenate([one.max_scores(i) for i in range(len(blocks))]))
    for i in range(len(blocks)):
        assert many.native_hss(i) == one.native_hss(i) == op.expected_hss(blks[i])
    one.close(); many.close()
