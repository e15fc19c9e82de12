"""ctypes binding of oracle/liboracle.so (the CPU restatement; TEST INFRASTRUCTURE, never shipped)."""
import ctypes as C
import gzip
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORC, "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class orc_params(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("Delta", "Omega", "omega", "stopPenalty_0", "stopPenalty_k")]


class orc_hss(C.Structure):
    _fields_ = [("strand", C.c_int), ("frame", C.c_int), ("startSite", C.c_int), ("endSite", C.c_int), ("score", C.c_float)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp, i = C.c_void_p, C.c_int
        lib.orc_transcode.restype = C.POINTER(C.c_int)
        lib.orc_blosum62.restype = C.POINTER(C.c_int)
        lib.orc_seq_length.argtypes = [vp, i]
        lib.orc_score_aln.argtypes = [vp, i, i, vp, vp, vp, C.POINTER(orc_params), C.POINTER(orc_hss), i]
        lib.orc_score_strand.argtypes = [vp, i, i, vp, vp, C.POINTER(orc_params), i, C.POINTER(orc_hss), i, vp]
        lib.orc_sample_max.argtypes = [vp, vp, i, i, vp, vp, vp, C.POINTER(orc_params)]
        lib.orc_sample_max.restype = C.c_double
        lib.orc_pair_row.argtypes = [vp, i, i, vp, vp, C.POINTER(orc_params), i, vp]
        lib.orc_rev_aln.argtypes = [vp, i, i, vp]
        lib.orc_calculate_bg.argtypes = [C.c_float, vp, C.c_float, vp, vp, vp]
        lib.orc_count_freqs.argtypes = [vp, i, i, vp]
        lib.orc_cells.argtypes = [i, i]
        lib.orc_cells.restype = C.c_double
        self.blosum62 = np.array([lib.orc_blosum62()[k] for k in range(576)], dtype=np.int32)
        self.transcode = np.array([lib.orc_transcode()[k] for k in range(64)], dtype=np.int32)

    @staticmethod
    def params(Delta=-10.0, Omega=-4.0, omega=-2.0, stopPenalty_0=-9999.0, stopPenalty_k=-8.0):
        return orc_params(Delta, Omega, omega, stopPenalty_0, stopPenalty_k)

    def score_aln(self, rows, scores_fwd, scores_rev, params, blosum=None):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        N, cols = rows.shape
        sf = np.ascontiguousarray(scores_fwd, dtype=np.float32)
        sr = np.ascontiguousarray(scores_rev, dtype=np.float32)
        bl = np.ascontiguousarray(self.blosum62 if blosum is None else blosum, dtype=np.int32)
        cap = 1024
        while True:
            out = (orc_hss * cap)()
            n = self.lib.orc_score_aln(rows.ctypes.data, N, cols, sf.ctypes.data, sr.ctypes.data, bl.ctypes.data,
                                       C.byref(params), out, cap)
            if n > cap:
                cap = n
                continue
            return [(chr(out[k].strand), out[k].frame, out[k].startSite, out[k].endSite, np.float32(out[k].score))
                    for k in range(n)]

    def sample_max(self, rows, sample, scores_fwd, scores_rev, params, blosum=None):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        sample = np.ascontiguousarray(sample, dtype=np.uint8)
        N, cols = rows.shape
        sf = np.ascontiguousarray(scores_fwd, dtype=np.float32)
        sr = np.ascontiguousarray(scores_rev, dtype=np.float32)
        bl = np.ascontiguousarray(self.blosum62 if blosum is None else blosum, dtype=np.int32)
        return self.lib.orc_sample_max(rows.ctypes.data, sample.ctypes.data, N, cols, sf.ctypes.data, sr.ctypes.data,
                                       bl.ctypes.data, C.byref(params))

    def sample_maxima(self, rows, samples, scores_fwd, scores_rev, params, blosum=None):
        return np.array([self.sample_max(rows, s, scores_fwd, scores_rev, params, blosum) for s in samples])

    def rev_aln(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        out = np.empty_like(rows)
        self.lib.orc_rev_aln(rows.ctypes.data, rows.shape[0], rows.shape[1], out.ctypes.data)
        return out

    def dense_S(self, rows, scores, params, blosum=None):
        """S[b][i] of one strand (getMultipleScoreMatrix, src/score.c:811-848) as orc_score_strand materialises it for small
        L: array [L+1][L+1], entries outside b = 1.., i = b+2, b+5, ... are 0."""
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        N, cols = rows.shape
        L = self.lib.orc_seq_length(rows.ctypes.data, cols)
        sc = np.ascontiguousarray(scores, dtype=np.float32)
        bl = np.ascontiguousarray(self.blosum62 if blosum is None else blosum, dtype=np.int32)
        S = np.zeros((L + 1, L + 1), dtype=np.float32)
        out = (orc_hss * 4096)()
        self.lib.orc_score_strand(rows.ctypes.data, N, cols, sc.ctypes.data, bl.ctypes.data, C.byref(params), ord("+"), out, 4096,
                                  S.ctypes.data)
        return S

    def pair_row(self, rows, scores, params, b, blosum=None):
        """Sk[k][state][b][0..L] of one strand (rows already in that strand's orientation): array [N][3][L+1]."""
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        N, cols = rows.shape
        L = self.lib.orc_seq_length(rows.ctypes.data, cols)
        sc = np.ascontiguousarray(scores, dtype=np.float32)
        bl = np.ascontiguousarray(self.blosum62 if blosum is None else blosum, dtype=np.int32)
        out = np.zeros((N, 3, L + 1), dtype=np.float32)
        self.lib.orc_pair_row(rows.ctypes.data, N, cols, sc.ctypes.data, bl.ctypes.data, C.byref(params), b, out.ctypes.data)
        return out

    def calculate_bg(self, dist, freqs, kappa, blosum=None):
        fr = np.ascontiguousarray(freqs, dtype=np.float32)
        bl = np.ascontiguousarray(self.blosum62 if blosum is None else blosum, dtype=np.int32)
        sc = np.zeros(4, dtype=np.float32)
        pr = np.zeros(4, dtype=np.float32)
        self.lib.orc_calculate_bg(C.c_float(dist), fr.ctypes.data, C.c_float(kappa), bl.ctypes.data, sc.ctypes.data,
                                  pr.ctypes.data)
        return sc

    def count_freqs(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        f = np.zeros(4, dtype=np.float32)
        self.lib.orc_count_freqs(rows.ctypes.data, rows.shape[0], rows.shape[1], f.ctypes.data)
        return f


_oracle = None


def build():
    src = [os.path.join(ORC, f) for f in ("rnacode_oracle.c", "rnacode_oracle.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", ORC, "liboracle"], check=True, stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _oracle
    if _oracle is None:
        _oracle = Oracle(C.CDLL(build()))
    return _oracle


def golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rb") as fh:
        return json.loads(fh.read())


GOLDEN_SETS = ["coding_aln", "noncoding_aln", "coding_maf", "noncoding_maf", "genomic_maf", "genomic_pre_maf",
               "coding_aln_pars", "synth_gappy", "synth_gapfree"]


def rows_array(strings):
    return np.frombuffer("".join(strings).encode("latin-1"), dtype=np.uint8).reshape(len(strings), -1)


def block_arrays(doc, blk):
    """(rows, scores_fwd, scores_rev, samples or None) of one golden block."""
    rows = rows_array(blk["rows"])
    sf = np.array(blk["scores_fwd"], dtype=np.float32)
    sr = np.array(blk["scores_rev"], dtype=np.float32)
    smp = None
    if blk.get("samples"):
        smp = np.stack([rows_array(s) for s in blk["samples"]])
    return rows, sf, sr, smp


def golden_params(doc):
    p = doc["params"]
    return dict(Delta=p["Delta"], Omega=p["Omega"], omega=p["omega"], stopPenalty_0=p["stopPenalty_0"],
                stopPenalty_k=p["stopPenalty_k"])


def expected_hss(blk):
    return [(h["strand"], h["frame"], h["startSite"], h["endSite"], np.float32(h["score"])) for h in blk["native_hss"]]
