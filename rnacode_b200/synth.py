"""Synthetic alignment generator: seeded, a pure function of (seed, block index).

root row uniform over ACGT; species s = copy of the root (even s) or of the previous row (odd s) with
i.i.d. substitutions at rate r_s ~ U(0.02, 0.32) to a uniform base; gap runs of length in {1,2,3,3,6}
started with probability 0.0067 per position in every row.  Null alignments ("samples") for benchmarks
are drawn the same way without gaps (the library re-imposes the native gap pattern, as
reintroduceGaps does, src/misc.c:127-148).
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
GAP = ord("-")
GAP_RUNS = np.array([1, 2, 3, 3, 6])


def _rng(seed, index, stream):
    return np.random.default_rng([int(seed), int(index), int(stream)])


def rates(seed, index, N):
    return _rng(seed, index, 0).uniform(0.02, 0.32, size=N)


def _mutate(rng, src, rate):
    out = src.copy()
    hit = rng.random(src.shape) < rate
    out[hit] = ACGT[rng.integers(0, 4, size=int(hit.sum()))]
    return out


def synth_block(seed, index, N, cols, gap_rate=0.0067):
    """Native alignment block: uint8 array (N, cols) over 'ACGT-'."""
    rng = _rng(seed, index, 1)
    r = rates(seed, index, N)
    rows = np.empty((N, cols), dtype=np.uint8)
    rows[0] = ACGT[rng.integers(0, 4, size=cols)]
    for s in range(1, N):
        src = rows[0] if s % 2 == 0 else rows[s - 1]
        rows[s] = _mutate(rng, src, r[s])
    if gap_rate > 0:
        for s in range(N):
            starts = np.nonzero(rng.random(cols) < gap_rate)[0]
            for st in starts:
                ln = GAP_RUNS[rng.integers(0, len(GAP_RUNS))]
                rows[s, st:st + ln] = GAP
        # keep at least 3 reference positions
        if (rows[0] != GAP).sum() < 3:
            rows[0, :min(3, cols)] = ACGT[:min(3, cols)]
    return rows


def synth_samples(seed, index, n, N, cols):
    """n null alignments (n, N, cols), gap-free, each evolved star-like from its own random root."""
    rng = _rng(seed, index, 2)
    r = rates(seed, index, N)
    out = np.empty((n, N, cols), dtype=np.uint8)
    root = rng.integers(0, 4, size=(n, cols), dtype=np.uint8)
    out[:, 0, :] = ACGT[root]
    for s in range(1, N):
        hit = rng.random((n, cols)) < r[s]
        repl = rng.integers(0, 4, size=(n, cols), dtype=np.uint8)
        out[:, s, :] = ACGT[np.where(hit, repl, root)]
    return out


def synth_tree(seed, index, N):
    """A star tree for the on-GPU simulation (rc_tree_desc arrays: parent, row, cum): root = node 0, one tip per
    alignment row; branch k mutates a site with the block's rate r_k to a uniform base (Jukes-Cantor-like), root
    frequencies uniform.  Same null model as synth_samples, drawn on the device instead of the host."""
    r = rates(seed, index, N)
    parent = np.array([-1] + [0] * N, dtype=np.int32)
    row = np.array([-1] + list(range(N)), dtype=np.int32)
    cum = np.zeros((N + 1, 16), dtype=np.float64)
    cum[0, :4] = [0.25, 0.5, 0.75, 1.0]
    for k in range(N):
        P = np.full((4, 4), r[k] / 4.0)
        P[np.arange(4), np.arange(4)] += 1.0 - r[k]
        cum[k + 1] = np.cumsum(P, axis=1).reshape(16)
    return parent, row, cum


def synth_scores(seed, index, N):
    """Plausible expected-score tables (bgModel.scores, src/score.h:35) without running a tree:
    the values calculateBG produces shrink with the reference-species distance; we draw a distance per
    species and interpolate between the near and far regimes observed on the examples."""
    rng = _rng(seed, index, 3)
    d = rng.uniform(0.02, 0.6, size=N)
    near = np.array([5.32, 1.01, -0.92, -1.61])
    far = np.array([5.25, 0.55, -1.05, -1.40])
    w = (d / 0.6)[:, None]
    fwd = (near * (1 - w) + far * w + rng.normal(0, 0.01, size=(N, 4))).astype(np.float32)
    rev = (fwd + rng.normal(0, 0.01, size=(N, 4))).astype(np.float32)
    return fwd, rev


def to_maf(blocks, path):
    """Write blocks (list of uint8 arrays) as MAF, naming rows like SURVEY 8(d)."""
    with open(path, "w") as fh:
        fh.write("##maf version=1\n")
        for bi, rows in enumerate(blocks):
            fh.write("a score=0\n")
            for s in range(rows.shape[0]):
                seq = rows[s].tobytes().decode()
                ung = len(seq) - seq.count("-")
                fh.write("s sp%d.chr1 %d %d + 100000000 %s\n" % (s, 1000 + bi * 10000, ung, seq))
            fh.write("\n")


def ungapped_len(rows):
    return int((rows[0] != GAP).sum())


def P_of_L(L):
    """Number of (start, end) pairs: sum over frames of sites*(sites+1)/2."""
    return sum(((L - f) // 3) * ((L - f) // 3 + 1) // 2 for f in range(3)) if L >= 3 else 0


def cells(N, L, n_samples):
    """Algorithmic DP cells: (n+1) * 2 * (N-1) * P(L)."""
    return (n_samples + 1) * 2 * (N - 1) * P_of_L(L)
