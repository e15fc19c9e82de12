"""GPU parity tests proper: the CUDA path, called through the C ABI, against (1) golden dumps of the
unmodified reference and (2) the CPU oracle on seeded synthetic inputs.  Bit-exact: HSS strand / frame /
codon coordinates, float32 scores, per-sample maxima."""
import numpy as np
import pytest

from tests import oracle_py as op

pytestmark = pytest.mark.gpu


def _capi():
    from rnacode_b200 import capi
    return capi


def _block(rows, sf, sr, smp=None):
    return _capi().Block(rows, sf, sr, smp)


@pytest.mark.parametrize("name", op.GOLDEN_SETS)
def test_golden_native_and_samples(rc_ctx, name):
    capi = _capi()
    doc = op.golden(name)
    prm = capi.make_params(**op.golden_params(doc))
    blocks, blks = [], []
    for blk in doc["blocks"]:
        if blk.get("skipped"):
            continue
        rows, sf, sr, smp = op.block_arrays(doc, blk)
        blocks.append(_block(rows, sf, sr, smp))
        blks.append(blk)
    bt = rc_ctx.batch(blocks, prm, doc["blosum"])
    bt.upload()
    bt.run()
    bt.download()
    for i, blk in enumerate(blks):
        assert bt.native_hss(i) == op.expected_hss(blk), (name, blk["index"], "native HSS")
        if blocks[i].n_samples:
            got = bt.max_scores(i).astype(np.float32)
            exp = np.array(blk["maxScores"][:blocks[i].n_samples], dtype=np.float32)
            assert np.array_equal(got, exp), (name, blk["index"], "sample maxima")
    assert bt.stats()["launches"] > 0
    bt.close()


@pytest.mark.parametrize("name", ["coding_aln", "synth_gappy", "genomic_pre_maf"])
def test_one_block_api_matches_reference(rc_ctx, name):
    """rc_score_aln / rc_score_samples (the scoreAln-shaped calls)."""
    capi = _capi()
    doc = op.golden(name)
    prm = capi.make_params(**op.golden_params(doc))
    for blk in doc["blocks"][:4]:
        if blk.get("skipped"):
            continue
        rows, sf, sr, smp = op.block_arrays(doc, blk)
        b = _block(rows, sf, sr, smp)
        assert rc_ctx.score_aln(b, prm, doc["blosum"]) == op.expected_hss(blk)
        if smp is not None:
            got = rc_ctx.score_samples(b, prm, doc["blosum"]).astype(np.float32)
            assert np.array_equal(got, np.array(blk["maxScores"][:len(smp)], dtype=np.float32))


@pytest.mark.parametrize("mode", ["force_dense", "band1"])
def test_dense_fallback_is_exact(mode, oracle):
    """The exact dense-S path (taken on tie-band overflow) gives the same answers as the digest path."""
    capi = _capi()
    ctx = capi.Context(0)
    if mode == "force_dense":
        ctx.set_option("force_dense", 1)
    else:
        ctx.set_option("band_slots", 1)
    try:
        for name in ("coding_aln", "synth_gappy"):
            doc = op.golden(name)
            prm = capi.make_params(**op.golden_params(doc))
            for blk in doc["blocks"]:
                if blk.get("skipped"):
                    continue
                rows, sf, sr, smp = op.block_arrays(doc, blk)
                smp = smp[:6] if smp is not None else None
                b = _block(rows, sf, sr, smp)
                bt = ctx.batch([b], prm, doc["blosum"])
                bt.upload(); bt.run(); bt.download()
                assert bt.native_hss(0) == op.expected_hss(blk), (mode, name, blk["index"])
                if smp is not None:
                    assert np.array_equal(bt.max_scores(0).astype(np.float32),
                                          np.array(blk["maxScores"][:len(smp)], dtype=np.float32))
                if mode == "force_dense":
                    assert bt.stats()["dense_fallbacks"] == 1 + b.n_samples
                bt.close()
    finally:
        ctx.close()


SHAPES = [
    # (N, cols, n_samples, gap_rate)
    (3, 9, 4, 0.0), (3, 3, 2, 0.0), (4, 5, 2, 0.0), (2, 60, 3, 0.02), (10, 120, 33, 0.0067), (10, 120, 5, 0.0),
    (6, 47, 7, 0.05), (6, 48, 7, 0.05), (6, 49, 7, 0.05), (10, 97, 3, 0.1), (26, 150, 5, 0.02), (50, 130, 3, 0.02),
    (100, 96, 2, 0.02), (10, 600, 4, 0.02), (8, 1000, 2, 0.0067),
    # wide alignments (k_dp_chain: species chunks pipelined through the warps of a CTA): 2, 3, 4, 9 and 16 chunks,
    # chunk sizes with and without a dummy species, rows longer than several hand-off stages; 17 chunks -> k_dp
    (18, 700, 3, 0.02), (25, 520, 2, 0.03), (26, 333, 2, 0.0), (38, 410, 2, 0.03), (100, 260, 1, 0.02),
    (193, 120, 1, 0.02), (200, 70, 1, 0.02),
    # beyond 16 chunks the shared-memory DP takes over: fewer warps per CTA, and a single-stage ring at the
    # reference's maximum of 500 rows (MAX_NUM_NAMES, src/rnaz_utils.h:7)
    (300, 50, 1, 0.02), (500, 40, 1, 0.02),
    # wide alignments in short blocks with >= 16 instances: sample-major DP, one launch per species chunk of 2-3 quads
    # (layout 5), partial species sums handed over in global memory: 2, 3, 7, 9 and 25 chunks, with dummy species
    (18, 200, 17, 0.0), (26, 150, 40, 0.02), (37, 90, 64, 0.05), (50, 130, 33, 0.02), (100, 96, 20, 0.03), (101, 70, 31, 0.02),
    (300, 45, 15, 0.02),
    # the same with frames too long for a resident sigma table: streamed in segments (k_dp_smps), plain and chunked
    (10, 600, 16, 0.02), (8, 1000, 15, 0.0067), (5, 1201, 31, 0.01), (26, 520, 15, 0.03), (18, 700, 17, 0.02), (50, 450, 16, 0.02),
    # reference rows so gappy that a chunk of 216 positions spans more than 256 columns: k_sigma_smp gathers bytes instead of
    # staging a window of raw columns
    (6, 700, 17, 0.08), (10, 900, 16, 0.12),
    # 13-16 scored species: row-major blocks take k_dp_chain with two warps of 6-8 species (k_dp_reg<13..16> spills),
    # blocks with >= 16 instances stay on the one-launch sample-major kernels
    (14, 700, 3, 0.02), (15, 2000, 1, 0.01), (16, 450, 2, 0.03), (17, 600, 2, 0.02), (17, 1500, 20, 0.01),
    (14, 150, 20, 0.02), (17, 150, 33, 0.02), (15, 900, 16, 0.02),
    # BASELINE config 5's row count (100-way): k_dp_chain with 9 chunks in two passes on rows of up to 330 codons, with few
    # and with >= 16 instances; short 100-way blocks take the chunked sample-major route
    (100, 1000, 2, 0.0067), (100, 1000, 17, 0.02), (100, 200, 33, 0.0067), (100, 60, 40, 0.0067),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "N%d_c%d_n%d_g%g" % s)
def test_synthetic_vs_oracle(rc_ctx, oracle, shape):
    from rnacode_b200 import synth
    capi = _capi()
    N, cols, n, gr = shape
    prm = capi.make_params()
    oprm = oracle.params()
    blocks, data = [], []
    for idx in range(3):
        rows = synth.synth_block(21, idx * 1000 + N * 7 + cols, N, cols, gap_rate=gr)
        sf, sr = synth.synth_scores(21, idx, N)
        smp = synth.synth_samples(21, idx * 1000 + cols, n, N, cols)
        blocks.append(_block(rows, sf, sr, smp))
        data.append((rows, sf, sr, smp))
    bt = rc_ctx.batch(blocks, prm, oracle.blosum62)
    bt.upload(); bt.run(); bt.download()
    for i, (rows, sf, sr, smp) in enumerate(data):
        assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oprm), (shape, i)
        exp = oracle.sample_maxima(rows, smp, sf, sr, oprm).astype(np.float32)
        assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), (shape, i)
    bt.close()


@pytest.mark.parametrize("n", [2, 17], ids=lambda n: "n%d" % n)
def test_config4_shape_vs_oracle(rc_ctx, oracle, n):
    """BASELINE config 4's block shape, 50 species x 5000 columns (4.1e8 DP cells per alignment): k_dp_chain with W = 5 warps
    over rows of up to 1666 codons (multi-stage hand-off, chain_tasks), native HSS list and the maxima of n null alignments
    against the streaming oracle, bit for bit (src/score.c:496-535, :830-845, :864-974)."""
    from rnacode_b200 import synth
    capi = _capi()
    N, cols = 50, 5000
    rows = synth.synth_block(2, 4000 + n, N, cols, gap_rate=0.0067)
    sf, sr = synth.synth_scores(2, n, N)
    smp = synth.synth_samples(2, 4000 + n, n, N, cols)
    bt = rc_ctx.batch([_block(rows, sf, sr, smp)], capi.make_params(), oracle.blosum62)
    bt.upload(); bt.run(); bt.download()
    hss = bt.native_hss(0)
    assert len(hss) > 50
    assert hss == oracle.score_aln(rows, sf, sr, oracle.params())
    exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params()).astype(np.float32)
    assert np.array_equal(bt.max_scores(0).astype(np.float32), exp)
    bt.close()


def test_config5_mixed_lengths_vs_oracle(rc_ctx, oracle):
    """BASELINE config 5's regime: 100-way blocks of mixed length (log-uniform 60..2000 columns) in ONE batch -- the routes
    change with the length (chunked sample-major, chain in passes) -- with 33 null alignments each, against the oracle."""
    from rnacode_b200 import synth
    capi = _capi()
    rng = np.random.default_rng(5)
    lens = [60, 2000] + [int(round(float(np.exp(rng.uniform(np.log(60), np.log(2000)))))) for _ in range(6)]
    blocks, data = [], []
    for idx, cols in enumerate(lens):
        rows = synth.synth_block(3, idx, 100, cols, gap_rate=0.0067)
        sf, sr = synth.synth_scores(3, idx, 100)
        smp = synth.synth_samples(3, idx, 33, 100, cols)
        blocks.append(_block(rows, sf, sr, smp))
        data.append((rows, sf, sr, smp))
    bt = rc_ctx.batch(blocks, capi.make_params(), oracle.blosum62)
    bt.upload(); bt.run(); bt.download()
    for i, (rows, sf, sr, smp) in enumerate(data):
        assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oracle.params()), (i, lens[i])
        exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params()).astype(np.float32)
        assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), (i, lens[i])
    bt.close()


@pytest.mark.parametrize("opts", [{"no_fused": 0}, {"no_fused": 1}, {"no_fold": 1}, {"tail_max": 12}, {"no_fused": 0, "tail_max": 12},
                                  {"no_sig_p2": 1}, {"no_fused": 1, "no_sig_p2": 1}, {"no_allf": 1}],
                         ids=lambda o: "+".join("%s%d" % kv for kv in sorted(o.items())))
def test_optional_sample_major_routes_vs_oracle(oracle, opts):
    """The routes of the sample-major family, each switched on and off: k_dp_smpf (the DP CTA builds its sigma table itself from
    the packed rows of k_pack2; default where it keeps two CTAs per SM) against k_sigma_p2 / k_sigma_smp + k_dp_smp (sigma tables
    in HBM, built from the packed rows or -- no_sig_p2 -- from class bytes), its folded last group
    (several start-codon pairs side by side when at most 16 instances are left), and the tail split (the last 1..12 instances
    of a block scored row-major; off by default).  Same answers as the oracle, bit for bit."""
    from rnacode_b200 import synth
    capi = _capi()
    ctx = capi.Context(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        # instance counts that leave 1, 5, 8, 13 and 16 instances in the last group (folded: 32, 4, 4, 2 and 2 pairs side by side)
        shapes = [(10, 120, 100), (10, 121, 69), (4, 45, 70), (17, 150, 33), (6, 250, 40), (26, 150, 40), (100, 96, 36), (50, 130, 75),
                  (10, 120, 32), (9, 118, 44), (12, 33, 79), (3, 7, 47), (5, 200, 16)]
        blocks, data = [], []
        for idx, (N, cols, n) in enumerate(shapes):
            rows = synth.synth_block(8, idx, N, cols, gap_rate=0.02)
            sf, sr = synth.synth_scores(8, idx, N)
            smp = synth.synth_samples(8, idx, n, N, cols)
            blocks.append(_block(rows, sf, sr, smp))
            data.append((rows, sf, sr, smp))
        bt = ctx.batch(blocks, capi.make_params(), oracle.blosum62)
        bt.upload(); bt.run(); bt.download()
        for i, (rows, sf, sr, smp) in enumerate(data):
            assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oracle.params()), (opts, shapes[i])
            exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params()).astype(np.float32)
            assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), (opts, shapes[i])
        bt.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("omega", [-2.0, -4.0, -0.5, -1.5], ids=lambda w: "omega%g" % w)
@pytest.mark.parametrize("reg_tu", [1, 0, -1], ids=lambda v: "reg_tu%d" % v)
def test_three_addition_kernel_vs_oracle(oracle, reg_tu, omega):
    """k_dp_regtu carries max(S1, S2) instead of S1 and S2 and brings the smaller one up to date at a frameshift -- n deferred
    additions of omega collapsed into one FMA per binade (needs omega = -2^k; -1.5 must fall back to k_dp_reg).  Forced on,
    forced off and on its own choice, on row-major blocks from gap-free to very gappy (several frameshifts per species and
    row, rows that start between two frameshifts, sums that cross many binades): the oracle's answers, bit for bit."""
    from rnacode_b200 import synth
    capi = _capi()
    ctx = capi.Context(0)
    ctx.set_option("reg_tu", reg_tu)
    kw = dict(omega=omega)
    try:
        shapes = [(8, 1000, 2, 0.0067), (10, 600, 4, 0.02), (10, 2400, 2, 0.0005), (5, 900, 2, 0.08), (12, 700, 3, 0.02),
                  (3, 300, 2, 0.05), (10, 1500, 1, 0.0), (2, 2000, 3, 0.03), (13, 500, 2, 0.01)]
        blocks, data = [], []
        for idx, (N, cols, n, gr) in enumerate(shapes):
            rows = synth.synth_block(77, idx, N, cols, gap_rate=gr)
            sf, sr = synth.synth_scores(77, idx, N)
            smp = synth.synth_samples(77, idx, n, N, cols)
            blocks.append(_block(rows, sf, sr, smp))
            data.append((rows, sf, sr, smp))
        bt = ctx.batch(blocks, capi.make_params(**kw), oracle.blosum62)
        bt.upload(); bt.run(); bt.download()
        for i, (rows, sf, sr, smp) in enumerate(data):
            assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oracle.params(**kw)), (reg_tu, omega, shapes[i])
            exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params(**kw)).astype(np.float32)
            assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), (reg_tu, omega, shapes[i])
        bt.close()
    finally:
        ctx.close()


def test_near_ties_after_the_row_maximum(rc_ctx, oracle):
    """The species-sum fold of the sample-major kernels treats a positive sum just below a row's maximum exactly (folds_exact).
    Alignments made of a few repeated columns give rows full of exact ties and of near ties (sums that differ by rounding
    only); gap-free and gappy, short and streamed frames."""
    from rnacode_b200 import synth
    capi = _capi()
    rng = np.random.default_rng(99)
    blocks, data = [], []
    for idx, (N, cols, n) in enumerate([(6, 90, 40), (10, 120, 33), (10, 600, 20), (4, 300, 64)]):
        base = synth.synth_block(9, idx, N, 9, gap_rate=0.0)          # nine columns, tiled: periodic sigma, many equal sums
        rows = np.tile(base, (1, cols // 9 + 1))[:, :cols].copy()
        if idx % 2:
            rows[1, 30:32] = synth.GAP
        sf, sr = synth.synth_scores(9, idx, N)
        sf[:, 1:] = np.round(sf[:, 1:] * 4) / 4                       # quarter-valued expected scores: exact cancellations
        sr[:, 1:] = np.round(sr[:, 1:] * 4) / 4
        smp = np.stack([np.tile(synth.synth_block(10 + s, idx, N, 9, gap_rate=0.0), (1, cols // 9 + 1))[:, :cols] for s in range(n)])
        blocks.append(_block(rows, sf, sr, smp))
        data.append((rows, sf, sr, smp))
    bt = rc_ctx.batch(blocks, capi.make_params(), oracle.blosum62)
    bt.upload(); bt.run(); bt.download()
    for i, (rows, sf, sr, smp) in enumerate(data):
        assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oracle.params()), i
        exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params()).astype(np.float32)
        assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), i
    bt.close()


def test_mixed_batch_and_chunking(oracle):
    """Blocks of different shapes in one batch, with a scratch budget so small that instances of one block
    are split across chunks."""
    from rnacode_b200 import synth
    capi = _capi()
    ctx = capi.Context(0)
    ctx.set_option("scratch_mb", 1)
    try:
        shapes = [(10, 200, 40), (4, 30, 3), (30, 90, 10), (3, 2, 2), (12, 333, 9),
                  # sample-major layouts whose instance groups of 32 are split over several items / chunks:
                  # chunked wide (layout 5), streamed (k_dp_smps), chain in passes
                  (30, 90, 70), (10, 600, 40), (80, 300, 3)]
        blocks, data = [], []
        for idx, (N, cols, n) in enumerate(shapes):
            rows = synth.synth_block(5, idx, N, cols, gap_rate=0.02)
            sf, sr = synth.synth_scores(5, idx, N)
            smp = synth.synth_samples(5, idx, n, N, cols)
            blocks.append(_block(rows, sf, sr, smp))
            data.append((rows, sf, sr, smp))
        bt = ctx.batch(blocks, capi.make_params(), oracle.blosum62)
        bt.upload(); bt.run(); bt.download()
        for i, (rows, sf, sr, smp) in enumerate(data):
            if synth.ungapped_len(rows) < 3:
                assert bt.native_hss(i) == []
                assert np.all(bt.max_scores(i) == -1.0)
                continue
            assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oracle.params()), i
            exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params()).astype(np.float32)
            assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), i
        bt.close()
    finally:
        ctx.close()


def test_chain_kernel_equals_generic_kernel(oracle):
    """The chained-warp DP (layout 3) and the generic shared-memory DP (no_chain) agree bit for bit, and wide
    alignments with omega > 0 (dummy species not neutral) are routed to the generic kernel and still exact."""
    from rnacode_b200 import synth
    capi = _capi()
    rows = synth.synth_block(33, 2, 31, 900, gap_rate=0.03)
    sf, sr = synth.synth_scores(33, 2, 31)
    smp = synth.synth_samples(33, 2, 3, 31, 900)
    res = []
    for no_chain, no_rows3 in ((0, 0), (1, 0), (0, 1)):  # (0, 1): the chain kernel on sigma tiles written by the generic k_sigma
        ctx = capi.Context(0)
        ctx.set_option("no_chain", no_chain)
        ctx.set_option("no_sig_rows3", no_rows3)
        try:
            b = _block(rows, sf, sr, smp)
            res.append((ctx.score_aln(b, capi.make_params(), oracle.blosum62),
                        ctx.score_samples(b, capi.make_params(), oracle.blosum62).astype(np.float32)))
        finally:
            ctx.close()
    assert res[0][0] == res[1][0] == res[2][0] == oracle.score_aln(rows, sf, sr, oracle.params())
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][1], res[2][1])
    kw = dict(Delta=-6.0, Omega=-3.0, omega=0.25, stopPenalty_0=-100.0, stopPenalty_k=-5.0)
    ctx = capi.Context(0)
    try:
        b = _block(rows[:20, :300], sf[:20], sr[:20], smp[:, :20, :300])
        assert ctx.score_aln(b, capi.make_params(**kw), oracle.blosum62) == oracle.score_aln(rows[:20, :300], sf[:20], sr[:20],
                                                                                             oracle.params(**kw))
    finally:
        ctx.close()


def test_nondefault_parameters(rc_ctx, oracle):
    from rnacode_b200 import synth
    capi = _capi()
    kw = dict(Delta=-7.5, Omega=-3.25, omega=-1.5, stopPenalty_0=-50.0, stopPenalty_k=-6.0)
    rows = synth.synth_block(9, 1, 8, 150, gap_rate=0.04)
    sf, sr = synth.synth_scores(9, 1, 8)
    smp = synth.synth_samples(9, 1, 6, 8, 150)
    b = _block(rows, sf, sr, smp)
    assert rc_ctx.score_aln(b, capi.make_params(**kw), oracle.blosum62) == oracle.score_aln(rows, sf, sr, oracle.params(**kw))
    exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params(**kw)).astype(np.float32)
    assert np.array_equal(rc_ctx.score_samples(b, capi.make_params(**kw), oracle.blosum62).astype(np.float32), exp)


def test_bad_arguments(rc_ctx, oracle):
    capi = _capi()
    rows = np.frombuffer(b"ACGTACGTAC", dtype=np.uint8).reshape(1, 10)
    b = capi.Block(rows, np.zeros((1, 4)), np.zeros((1, 4)))
    with pytest.raises(capi.RcError):
        rc_ctx.score_aln(b, capi.make_params(), oracle.blosum62)


def test_random_alignments_vs_oracle(rc_ctx, oracle):
    """Seeded random sweep: shapes, gap rates, stray symbols (N, X, IUPAC, lower case), penalties and sample counts
    drawn at random; every block must match the oracle bit for bit (native HSS and per-sample maxima)."""
    from rnacode_b200 import synth
    capi = _capi()
    rng = np.random.default_rng(20261017)
    stray = np.frombuffer(b"NXRYnacgt", dtype=np.uint8)
    for rep in range(6):
        kw = dict(Delta=-float(rng.uniform(4, 14)), Omega=-float(rng.uniform(1, 6)), omega=-float(rng.uniform(0.5, 3)),
                  stopPenalty_0=-float(rng.uniform(50, 9999)), stopPenalty_k=-float(rng.uniform(2, 12)))
        blocks, data = [], []
        for idx in range(8):
            N = int(rng.choice([3, 4, 6, 9, 10, 13, 17, 18, 24, 33, 41]))
            cols = int(rng.integers(3, 420))
            n = int(rng.choice([1, 2, 5, 17, 33, 40]))
            rows = synth.synth_block(1000 + rep, idx, N, cols, gap_rate=float(rng.choice([0.0, 0.005, 0.02, 0.08])))
            hits = rng.random(rows.shape) < 0.004
            rows = rows.copy()
            rows[hits & (rows != synth.GAP)] = stray[rng.integers(0, len(stray), size=int((hits & (rows != synth.GAP)).sum()))]
            sf, sr = synth.synth_scores(1000 + rep, idx, N)
            smp = synth.synth_samples(1000 + rep, idx, n, N, cols)
            blocks.append(_block(rows, sf, sr, smp))
            data.append((rows, sf, sr, smp))
        bt = rc_ctx.batch(blocks, capi.make_params(**kw), oracle.blosum62)
        bt.upload(); bt.run(); bt.download()
        for i, (rows, sf, sr, smp) in enumerate(data):
            if synth.ungapped_len(rows) < 3:
                assert bt.native_hss(i) == []
                continue
            assert bt.native_hss(i) == oracle.score_aln(rows, sf, sr, oracle.params(**kw)), (rep, i, rows.shape)
            exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params(**kw)).astype(np.float32)
            assert np.array_equal(bt.max_scores(i).astype(np.float32), exp), (rep, i, rows.shape)
        bt.close()


@pytest.mark.parametrize("shape", [(8, 137, 0.02), (10, 600, 0.03), (3, 30, 0.0), (40, 210, 0.05), (2, 3, 0.0), (500, 40, 0.02)],
                         ids=lambda s: "N%d_c%d_g%g" % s)
def test_pair_rows_vs_oracle(rc_ctx, oracle, shape):
    """rc_pair_rows (the Sk_native rows backtrack() walks for --eps) against orc_pair_row, which is pinned to the
    reference's own matrices in tests/test_oracle_golden.py: both strands, first / inner / last rows, bit-exact."""
    from rnacode_b200 import synth
    capi = _capi()
    N, cols, gr = shape
    rows = synth.synth_block(33, N + cols, N, cols, gap_rate=gr)
    sf, sr = synth.synth_scores(33, 1, N)
    L = int((np.asarray(rows)[0] != ord("-")).sum())
    bs = sorted({1, 2, 3, max(1, L // 2), max(1, L - 3), max(1, L - 2), L})
    blk = _block(rows, sf, sr, None)
    for params in (dict(), dict(Delta=-9.5, Omega=-3.25, omega=-1.5, stopPenalty_0=-50.0)):
        prm = capi.make_params(**params)
        oprm = oracle.params(**params)
        rev = oracle.rev_aln(rows)
        for strand in (0, 1):
            got = rc_ctx.pair_rows(blk, prm, oracle.blosum62, strand, bs)
            assert got.shape == (len(bs), N, 3, L + 1)
            for r, b in enumerate(bs):
                exp = oracle.pair_row(rev if strand else rows, sr if strand else sf, oprm, b)
                assert np.array_equal(got[r].view(np.uint32), exp.view(np.uint32)), (shape, strand, b)


def test_pair_rows_rejects_bad_start(rc_ctx, oracle):
    from rnacode_b200 import synth
    capi = _capi()
    rows = synth.synth_block(33, 5, 4, 30, gap_rate=0.0)
    sf, sr = synth.synth_scores(33, 1, 4)
    blk = _block(rows, sf, sr, None)
    for b in (0, 31, -2):
        with pytest.raises(RuntimeError):
            rc_ctx.pair_rows(blk, capi.make_params(), oracle.blosum62, 0, [b])


def test_full_size_kernel_routes_agree(oracle):
    """BASELINE config 2 at full size (the 11 block shapes of examples/genomic.maf incl. 10 x 4806, -n 1000; the oracle would
    need hours): every block is scored by three different kernel routes -- the default (k_dp_reg for long frames, k_dp_smp /
    k_dp_smps for short and mid ones), everything row-major (no_smp), and everything streamed sample-major (k_dp_smps) --
    whose sigma layouts, task shapes and getHSS folds differ, with the null alignments drawn on the GPU from the same
    seeds.  All native HSS and all 11 x 1000 sample maxima must agree bit for bit."""
    import bench
    from rnacode_b200 import synth
    capi = _capi()
    blocks_np, n, seed, _ = bench.build_workload("genomic", 0)
    assert n == 1000 and max(r.shape[1] for r, _, _, _ in blocks_np) == 4806
    blocks = [capi.Block(rows, sf, sr, None, n_samples=n) for rows, sf, sr, _ in blocks_np]
    trees = [capi.Tree(*synth.synth_tree(seed, idx, rows.shape[0])) for rows, _, _, idx in blocks_np]
    seeds = [np.arange(1, n + 1, dtype=np.uint32) + 104729 * i for i in range(len(blocks))]
    results = []
    for opts in ({}, {"no_smp": 1}, {"smps_max_sites": 100000, "scratch_mb": 4096}):
        ctx = capi.Context(0)
        try:
            for k, v in opts.items():
                ctx.set_option(k, v)
            bt = ctx.batch(blocks, capi.make_params(), oracle.blosum62)
            for i in range(len(blocks)):
                bt.set_evolve(i, trees[i], seeds[i], capi.RC_RNG_MT19937)
            bt.upload(); bt.run(); bt.download()
            results.append(([bt.native_hss(i) for i in range(len(blocks))],
                            [bt.max_scores(i).copy() for i in range(len(blocks))], bt.stats()["dense_fallbacks"]))
            bt.close()
        finally:
            ctx.close()
    ref_hss, ref_max, _ = results[0]
    assert sum(len(h) for h in ref_hss) > 0 and all((m >= -1.0).all() for m in ref_max)
    for hss, mx, _ in results[1:]:
        assert hss == ref_hss
        for a, b in zip(mx, ref_max):
            assert np.array_equal(a, b)


def test_thread_scan_on_long_frames(oracle):
    """getHSS scan with one thread per (instance, strand, frame) -- taken by blocks with thousands of scans -- on frames of
    96..640 codons, forced here for a few instances (hss_thr_tasks = 1) so that the oracle can check it."""
    from rnacode_b200 import synth
    capi = _capi()
    ctx = capi.Context(0)
    ctx.set_option("hss_thr_tasks", 1)
    try:
        for (N, cols, n, gr) in [(10, 600, 5, 0.02), (6, 1900, 2, 0.01), (18, 700, 17, 0.02)]:
            rows = synth.synth_block(77, cols, N, cols, gap_rate=gr)
            sf, sr = synth.synth_scores(77, 3, N)
            smp = synth.synth_samples(77, cols, n, N, cols)
            bt = ctx.batch([_block(rows, sf, sr, smp)], capi.make_params(), oracle.blosum62)
            bt.upload(); bt.run(); bt.download()
            assert bt.native_hss(0) == oracle.score_aln(rows, sf, sr, oracle.params()), (N, cols)
            exp = oracle.sample_maxima(rows, smp, sf, sr, oracle.params()).astype(np.float32)
            assert np.array_equal(bt.max_scores(0).astype(np.float32), exp), (N, cols)
            bt.close()
    finally:
        ctx.close()
