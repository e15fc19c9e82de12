"""Pins the CPU oracle (oracle/rnacode_oracle.c) against dumps of the UNMODIFIED reference
(tests/golden/*.json.gz, produced by oracle/_ref/ref_probe -- see tests/golden/make_golden.py).
Bit-exact: HSS coordinates and float32 scores, per-sample maxima, background-model scores."""
import numpy as np
import pytest

from tests import oracle_py as op


@pytest.mark.parametrize("name", op.GOLDEN_SETS)
def test_tables_match_reference(oracle, name):
    doc = op.golden(name)
    assert list(oracle.transcode) == doc["transcode"]
    assert list(oracle.blosum62) == doc["blosum"]


@pytest.mark.parametrize("name", op.GOLDEN_SETS)
def test_native_hss_bit_exact(oracle, name):
    doc = op.golden(name)
    prm = oracle.params(**op.golden_params(doc))
    scored = 0
    for blk in doc["blocks"]:
        if blk.get("skipped"):
            continue
        if blk["N"] * blk["L"] ** 2 > 4e8:  # the 10x4806 block of genomic.maf: covered by the GPU suite, too slow here
            continue
        rows, sf, sr, _ = op.block_arrays(doc, blk)
        got = oracle.score_aln(rows, sf, sr, prm)
        assert got == op.expected_hss(blk), (name, blk["index"])
        scored += 1
    assert scored > 0


@pytest.mark.parametrize("name", op.GOLDEN_SETS)
def test_sample_maxima_bit_exact(oracle, name):
    doc = op.golden(name)
    prm = oracle.params(**op.golden_params(doc))
    checked = 0
    for blk in doc["blocks"]:
        if blk.get("skipped") or blk["N"] * blk["L"] ** 2 > 4e8:
            continue
        rows, sf, sr, smp = op.block_arrays(doc, blk)
        if smp is None:
            continue
        got = oracle.sample_maxima(rows, smp, sf, sr, prm)
        exp = np.array(blk["maxScores"][:len(smp)])
        assert np.array_equal(got.astype(np.float32), exp.astype(np.float32)), (name, blk["index"])
        checked += len(smp)
    assert checked > 0


@pytest.mark.parametrize("name", ["coding_aln", "genomic_pre_maf", "synth_gappy"])
def test_background_model_bit_exact(oracle, name):
    """calculateBG / probHKY / countFreqsMono restatement vs models[].scores of the reference."""
    doc = op.golden(name)
    for blk in doc["blocks"]:
        if blk.get("skipped"):
            continue
        rows, sf, sr, _ = op.block_arrays(doc, blk)
        fr = oracle.count_freqs(rows)
        assert np.array_equal(fr, np.array(blk["freqs_fwd"], dtype=np.float32))
        for k in range(1, blk["N"]):
            sc = oracle.calculate_bg(blk["dist"][k], blk["freqs_fwd"], blk["kappa"])
            assert np.array_equal(sc, sf[k]), (name, blk["index"], k)
            sc = oracle.calculate_bg(blk["dist"][k], blk["freqs_rev"], blk["kappa"])
            assert np.array_equal(sc, sr[k]), (name, blk["index"], k)


@pytest.mark.parametrize("name", ["coding_aln", "synth_gappy"])
def test_pair_rows_bit_exact(oracle, name):
    """orc_pair_row against rows of the reference's Sk_native / Sk_native_rev (oracle/ref_probe.c --sk-rows), the
    matrices backtrack() walks for the --eps plots (src/score.c:558-797)."""
    doc = op.golden("sk_rows")[name]
    prm = oracle.params(**op.golden_params(doc))
    checked = 0
    for blk in doc["blocks"]:
        if not blk.get("sk_rows"):
            continue
        rows, sf, sr, _ = op.block_arrays(doc, blk)
        rev = oracle.rev_aln(rows)
        N, L = blk["N"], blk["L"]
        for rec in blk["sk_rows"]:
            b = rec["b"]
            got = oracle.pair_row(rev if rec["strand"] else rows, sr if rec["strand"] else sf, prm, b)
            exp = rec["v"]
            for k in range(1, N):
                for x in range(3):
                    e = np.array(exp[(k - 1) * 3 + x], dtype=np.float32)
                    assert np.array_equal(got[k, x, b - 1:L + 1:3], e), (blk["index"], rec["strand"], b, k, x)
                    checked += len(e)
    assert checked > 1000


def test_pair_rows_consistent_with_multiple_score_matrix(oracle):
    """S[b][i] = max(sum_k max3(Sk[k][0..2][b][i]), Delta) / (N-1) (src/score.c:830-845): the rows orc_pair_row
    produces must reproduce, bit for bit, the S matrix the (golden-pinned) streaming restatement materialises."""
    from rnacode_b200 import synth
    prm = oracle.params()
    for idx, (N, cols, gr) in enumerate([(5, 60, 0.02), (9, 99, 0.05), (3, 33, 0.0)]):
        rows = synth.synth_block(91, idx, N, cols, gap_rate=gr)
        sf, _ = synth.synth_scores(91, idx, N)
        S = oracle.dense_S(rows, sf, prm)
        L = S.shape[0] - 1
        for b in range(1, L - 1):
            row = oracle.pair_row(rows, sf, prm, b)
            for i in range(b + 2, L + 1, 3):
                tot = np.float32(0.0)
                for k in range(1, N):
                    tot = np.float32(tot + np.float32(row[k, :, i].max()))
                exp = np.float32(max(tot, np.float32(prm.Delta))) / np.float32(N - 1)
                assert np.float32(exp) == S[b, i], (idx, b, i)


def test_genomic_maf_n1000_golden(oracle):
    """BASELINE config 2 at its own -n 1000 (tests/golden/genomic_maf_n1000.json.gz: per-sample maxima of the unmodified
    reference for all 1000 null alignments of every block of examples/genomic.maf, incl. the 10 x 4806 one).  The oracle
    redraws null alignments from the dumped seeds / tree (orc_evolve = seq-gen's MT19937 + HKY walk) and scores them: every
    native HSS list, and the maxima of the first 40 (10 x 4806 block) or 250 (short blocks) null alignments, bit for bit."""
    import ctypes as C
    doc = op.golden("genomic_maf_n1000")
    prm = oracle.params(**op.golden_params(doc))
    oracle.lib.orc_evolve.argtypes = [C.c_ulong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_void_p]
    checked = 0
    for blk in doc["blocks"]:
        if blk.get("skipped"):
            continue
        rows, sf, sr, _ = op.block_arrays(doc, blk)
        assert oracle.score_aln(rows, sf, sr, prm) == op.expected_hss(blk), blk["index"]
        nodes = blk["evolve"]["nodes"]
        parent = np.array([n["parent"] for n in nodes], dtype=np.int32)
        row = np.array([n["row"] for n in nodes], dtype=np.int32)
        cum = np.array([n["cum"] for n in nodes], dtype=np.float64)
        af = np.array(blk["evolve"]["addFreq"], dtype=np.float64)
        N, cols = rows.shape
        k = 40 if cols > 2000 else 250
        smp = np.zeros((k, N, cols), dtype=np.uint8)
        for s in range(k):
            oracle.lib.orc_evolve(int(blk["seeds"][s]), len(nodes), parent.ctypes.data, row.ctypes.data, cum.ctypes.data,
                                  af.ctypes.data, N, cols, smp[s].ctypes.data)
        got = oracle.sample_maxima(rows, smp, sf, sr, prm).astype(np.float32)
        assert np.array_equal(got, np.array(blk["maxScores"][:k], dtype=np.float32)), blk["index"]
        checked += k
    assert checked == 40 + 9 * 250
