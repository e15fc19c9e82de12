"""ctypes binding of include/rnacode_cuda.h (test / benchmark plumbing, not the product).

Mirrors the C ABI one to one; see the header for the reference functions each call replaces
(scoreAln src/score.c:1067, the sampling loop of getExtremeValuePars src/score.c:1004-1044).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

RC_OK = 0


class RcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libRNAcode_cuda error %d: %s" % (code, msg))
        self.code = code


class rc_params(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("Delta", "Omega", "omega", "stopPenalty_0", "stopPenalty_k")]


class rc_hss(C.Structure):
    _fields_ = [("strand", C.c_int), ("frame", C.c_int), ("startSite", C.c_int), ("endSite", C.c_int), ("score", C.c_float)]


class rc_block_desc(C.Structure):
    _fields_ = [("N", C.c_int), ("cols", C.c_int), ("rows", C.c_void_p), ("scores_fwd", C.c_void_p),
                ("scores_rev", C.c_void_p), ("n_samples", C.c_int), ("samples", C.c_void_p)]


class rc_tree_desc(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("parent", C.c_void_p), ("row", C.c_void_p), ("cum", C.c_void_p)]


RC_RNG_MT19937, RC_RNG_PHILOX = 0, 1


class Tree:
    """Flattened tree in seq-gen's evolution order (see rc_tree_desc in the header)."""

    def __init__(self, parent, row, cum):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.row = np.ascontiguousarray(row, dtype=np.int32)
        self.cum = np.ascontiguousarray(cum, dtype=np.float64).reshape(len(self.parent), 16)

    def desc(self):
        return rc_tree_desc(len(self.parent), self.parent.ctypes.data, self.row.ctypes.data, self.cum.ctypes.data)


class rc_batch_stats(C.Structure):
    _fields_ = [("cells", C.c_double), ("launches", C.c_longlong), ("dense_fallbacks", C.c_longlong),
                ("ms_pack", C.c_float), ("ms_sigma", C.c_float), ("ms_dp", C.c_float), ("ms_hss", C.c_float),
                ("dp_launches", C.c_longlong), ("h2d_bytes", C.c_size_t), ("d2h_bytes", C.c_size_t),
                ("device_bytes", C.c_size_t), ("ms_pack_kernel", C.c_float), ("pack_chars", C.c_double)]


EXPORTS = [
    "rc_create", "rc_destroy", "rc_last_error", "rc_default_params", "rc_set_stream", "rc_set_option",
    "rc_score_aln", "rc_score_samples", "rc_batch_create", "rc_batch_upload", "rc_batch_run", "rc_batch_download",
    "rc_batch_native_hss", "rc_batch_max_scores", "rc_batch_destroy", "rc_batch_get_stats", "rc_version",
    "rc_calibrate_issue", "rc_batch_set_evolve", "rc_score_samples_evolve", "rc_batch_get_sample_rows", "rc_device_count",
    "rc_pair_rows", "rc_batch_set_evolve_many", "rc_batch_max_scores_all",
]

_lib = None


def load():
    """Load lib/libRNAcode_cuda.so.  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RNACODE_CUDA_LIB", _build.LIB)  # override: A/B-testing of differently built libraries
    if not os.path.exists(path):
        raise ImportError("libRNAcode_cuda.so is missing (%s): run `python -m rnacode_b200.build` "
                          "or __graft_entry__.build(); no CPU fallback exists" % path)
    lib = C.CDLL(path)
    vp, i = C.c_void_p, C.c_int
    lib.rc_create.argtypes = [C.POINTER(vp), i]
    lib.rc_destroy.argtypes = [vp]
    lib.rc_destroy.restype = None
    lib.rc_last_error.argtypes = [vp]
    lib.rc_last_error.restype = C.c_char_p
    lib.rc_default_params.argtypes = [C.POINTER(rc_params)]
    lib.rc_default_params.restype = None
    lib.rc_set_stream.argtypes = [vp, vp]
    lib.rc_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    lib.rc_score_aln.argtypes = [vp, C.POINTER(rc_block_desc), C.POINTER(rc_params), vp, C.POINTER(rc_hss), i, C.POINTER(i)]
    lib.rc_score_samples.argtypes = [vp, C.POINTER(rc_block_desc), C.POINTER(rc_params), vp, vp]
    lib.rc_pair_rows.argtypes = [vp, C.POINTER(rc_block_desc), C.POINTER(rc_params), vp, i, i, vp, vp]
    lib.rc_batch_create.argtypes = [vp, C.POINTER(rc_block_desc), i, C.POINTER(rc_params), vp, C.POINTER(vp)]
    for f in ("rc_batch_upload", "rc_batch_run", "rc_batch_download"):
        getattr(lib, f).argtypes = [vp]
    lib.rc_batch_native_hss.argtypes = [vp, i, C.POINTER(rc_hss), i, C.POINTER(i)]
    lib.rc_batch_max_scores.argtypes = [vp, i, vp]
    lib.rc_batch_destroy.argtypes = [vp]
    lib.rc_batch_destroy.restype = None
    lib.rc_batch_get_stats.argtypes = [vp, C.POINTER(rc_batch_stats)]
    lib.rc_version.restype = C.c_char_p
    lib.rc_calibrate_issue.argtypes = [vp, C.POINTER(C.c_double)]
    lib.rc_batch_set_evolve.argtypes = [vp, i, C.POINTER(rc_tree_desc), vp, i]
    lib.rc_score_samples_evolve.argtypes = [vp, C.POINTER(rc_block_desc), C.POINTER(rc_tree_desc), vp, i,
                                            C.POINTER(rc_params), vp, vp]
    lib.rc_batch_get_sample_rows.argtypes = [vp, i, i, vp]
    lib.rc_batch_set_evolve_many.argtypes = [vp, i, i, vp, vp, i]
    lib.rc_batch_max_scores_all.argtypes = [vp, vp, C.c_size_t]
    _lib = lib
    return lib


def default_params():
    p = rc_params()
    load().rc_default_params(C.byref(p))
    return p


def make_params(Delta=-10.0, Omega=-4.0, omega=-2.0, stopPenalty_0=-9999.0, stopPenalty_k=-8.0):
    return rc_params(Delta, Omega, omega, stopPenalty_0, stopPenalty_k)


class Block:
    """Host-side view of one alignment block (arrays are kept alive by this object)."""

    def __init__(self, rows, scores_fwd, scores_rev, samples=None, n_samples=None):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        assert rows.ndim == 2
        self.N, self.cols = rows.shape
        self.rows = rows
        self.scores_fwd = np.ascontiguousarray(scores_fwd, dtype=np.float32).reshape(self.N, 4)
        self.scores_rev = np.ascontiguousarray(scores_rev, dtype=np.float32).reshape(self.N, 4)
        if samples is not None:
            if not (isinstance(samples, np.ndarray) and samples.dtype == np.uint8 and samples.flags.c_contiguous):
                samples = np.ascontiguousarray(samples, dtype=np.uint8)
            assert samples.shape[1:] == (self.N, self.cols)
        self.samples = samples
        self.n_samples = (n_samples or 0) if samples is None else samples.shape[0]

    @staticmethod
    def from_strings(rows, scores_fwd, scores_rev, samples=None):
        r = np.frombuffer("".join(rows).encode("latin-1"), dtype=np.uint8).reshape(len(rows), -1)
        s = None
        if samples is not None and len(samples):
            s = np.frombuffer("".join("".join(x) for x in samples).encode("latin-1"), dtype=np.uint8)
            s = s.reshape(len(samples), len(rows), -1)
        return Block(r, scores_fwd, scores_rev, s)

    def desc(self):
        return rc_block_desc(self.N, self.cols, self.rows.ctypes.data, self.scores_fwd.ctypes.data,
                             self.scores_rev.ctypes.data, self.n_samples,
                             self.samples.ctypes.data if self.samples is not None else None)


def _blosum_arr(blosum):
    b = np.ascontiguousarray(blosum, dtype=np.int32).reshape(576)
    return b


class Context:
    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.rc_create(C.byref(self.h), device)
        if rc != RC_OK:
            raise RcError(rc, "rc_create failed (no usable CUDA device %d?)" % device)

    def _check(self, rc):
        if rc != RC_OK:
            raise RcError(rc, self.lib.rc_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.rc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self.lib.rc_set_option(self.h, key.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.rc_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def score_aln(self, block, params, blosum):
        """rc_score_aln: list of (strand, frame, startSite, endSite, score float32)."""
        d = block.desc()
        bl = _blosum_arr(blosum)
        cap = 256
        while True:
            out = (rc_hss * cap)()
            n = C.c_int()
            rc = self.lib.rc_score_aln(self.h, C.byref(d), C.byref(params), bl.ctypes.data, out, cap, C.byref(n))
            if rc == -4:  # RC_ERR_CAPACITY
                cap = n.value
                continue
            self._check(rc)
            return [(chr(out[i].strand), out[i].frame, out[i].startSite, out[i].endSite, np.float32(out[i].score))
                    for i in range(n.value)]

    def pair_rows(self, block, params, blosum, strand, b_list):
        """rc_pair_rows: Sk[k][state][b][0..L] of the native alignment for every b in b_list: array [len(b_list)][N][3][L+1]."""
        d = block.desc()
        bl = _blosum_arr(blosum)
        bs = np.ascontiguousarray(b_list, dtype=np.int32)
        L = int((np.asarray(block.rows)[0] != ord("-")).sum())
        out = np.full((len(bs), d.N, 3, L + 1), np.nan, dtype=np.float32)
        self._check(self.lib.rc_pair_rows(self.h, C.byref(d), C.byref(params), bl.ctypes.data, strand, len(bs), bs.ctypes.data,
                                          out.ctypes.data))
        return out

    def score_samples(self, block, params, blosum):
        d = block.desc()
        bl = _blosum_arr(blosum)
        res = np.zeros(block.n_samples, dtype=np.float64)
        self._check(self.lib.rc_score_samples(self.h, C.byref(d), C.byref(params), bl.ctypes.data, res.ctypes.data))
        return res

    def calibrate_issue(self):
        v = C.c_double()
        self._check(self.lib.rc_calibrate_issue(self.h, C.byref(v)))
        return v.value

    def batch(self, blocks, params, blosum, descs=None):
        return Batch(self, blocks, params, blosum, descs)


class Batch:
    def __init__(self, ctx, blocks, params, blosum, descs=None):
        self.ctx = ctx
        self.blocks = list(blocks)
        # descs: a ctypes array built earlier with Batch.block_descs(blocks) (thousands of blocks: not rebuilt per batch)
        self._descs = descs if descs is not None else Batch.block_descs(self.blocks)
        self._blosum = _blosum_arr(blosum)
        self.h = C.c_void_p()
        ctx._check(ctx.lib.rc_batch_create(ctx.h, self._descs, len(self.blocks), C.byref(params), self._blosum.ctypes.data,
                                           C.byref(self.h)))

    @staticmethod
    def block_descs(blocks):
        return (rc_block_desc * len(blocks))(*[b.desc() for b in blocks])

    def set_evolve(self, i, tree, seeds, rng=RC_RNG_MT19937):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        d = tree.desc()
        self._keep = getattr(self, "_keep", []) + [tree, seeds]
        self.ctx._check(self.ctx.lib.rc_batch_set_evolve(self.h, i, C.byref(d), seeds.ctypes.data, rng))

    @staticmethod
    def evolve_plan(trees, seeds):
        """ctypes arrays for set_evolve_many, built once for a list of trees / per-block seed arrays (kept alive by the plan)."""
        seeds = [np.ascontiguousarray(s, dtype=np.uint32) for s in seeds]
        descs = (rc_tree_desc * len(trees))(*[t.desc() for t in trees])
        ptrs = (C.c_void_p * len(seeds))(*[s.ctypes.data for s in seeds])
        return {"descs": descs, "ptrs": ptrs, "keep": (trees, seeds), "n": len(trees)}

    def set_evolve_many(self, plan, rng=RC_RNG_MT19937, first=0):
        self._keep = getattr(self, "_keep", []) + [plan]
        self.ctx._check(self.ctx.lib.rc_batch_set_evolve_many(self.h, first, plan["n"], plan["descs"], plan["ptrs"], rng))

    def max_scores_all(self):
        n = sum(b.n_samples for b in self.blocks)
        res = np.zeros(n, dtype=np.float64)
        self.ctx._check(self.ctx.lib.rc_batch_max_scores_all(self.h, res.ctypes.data, n))
        return res

    def sample_rows(self, i, sample):
        b = self.blocks[i]
        out = np.zeros((b.N, b.cols), dtype=np.uint8)
        self.ctx._check(self.ctx.lib.rc_batch_get_sample_rows(self.h, i, sample, out.ctypes.data))
        return out

    def upload(self):
        self.ctx._check(self.ctx.lib.rc_batch_upload(self.h))

    def run(self):
        self.ctx._check(self.ctx.lib.rc_batch_run(self.h))

    def download(self):
        self.ctx._check(self.ctx.lib.rc_batch_download(self.h))

    def native_hss(self, i):
        cap = 256
        while True:
            out = (rc_hss * cap)()
            n = C.c_int()
            rc = self.ctx.lib.rc_batch_native_hss(self.h, i, out, cap, C.byref(n))
            if rc == -4:
                cap = n.value
                continue
            self.ctx._check(rc)
            return [(chr(out[k].strand), out[k].frame, out[k].startSite, out[k].endSite, np.float32(out[k].score))
                    for k in range(n.value)]

    def max_scores(self, i):
        res = np.zeros(self.blocks[i].n_samples, dtype=np.float64)
        self.ctx._check(self.ctx.lib.rc_batch_max_scores(self.h, i, res.ctypes.data))
        return res

    def stats(self):
        s = rc_batch_stats()
        self.ctx._check(self.ctx.lib.rc_batch_get_stats(self.h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in rc_batch_stats._fields_}

    def close(self):
        if self.h:
            self.ctx.lib.rc_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
