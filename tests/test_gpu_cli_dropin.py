"""Drop-in check of the whole program: the reference CLI with scoreAln / getExtremeValuePars replaced by
libRNAcode_cuda (integration/rnacode_cuda_shim.c, linked as oracle/_ref/RNAcode_cuda_det) must print what the
unmodified reference prints (tests/golden/cli_outputs.json.gz, produced by oracle/_ref/RNAcode_det with the
same deterministic seeds): identical HSS, coordinates, scores and p-values in default, --gtf and --tabular
format, with -b, -r, -s/-p, -c and -n."""
import os
import re
import subprocess

import pytest

from tests import oracle_py as op

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
CLI = os.path.join(REFDIR, "RNAcode_cuda_det")
EXAMPLES = os.path.join(REFDIR, "examples")


def _input(fname, tmp_path):
    """Path of a golden case's input: an example of the reference, or the synthetic MAF of mixed shapes (rebuilt here with
    the generator that produced the golden, tests/golden/make_golden.py:mixed_blocks)."""
    if fname == "synthetic:mixed":
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden
        from rnacode_b200 import synth
        p = os.path.join(str(tmp_path), "synth_mixed.maf")
        synth.to_maf(make_golden.mixed_blocks(), p)
        return p
    if fname == "synthetic:dupnames":
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden
        return make_golden.dupnames_maf(EXAMPLES, os.path.join(str(tmp_path), "dupnames.maf"))
    return os.path.join(EXAMPLES, fname)


def _norm(txt):
    # the footer of the default format reports CPU seconds
    return re.sub(r"scored in [0-9.]+ seconds", "scored in X seconds", txt)


CASES = sorted(op.golden("cli_outputs").keys())
PIPELINE = os.path.join(REFDIR, "RNAcode_b200_det")


# default: the null alignments are simulated on the GPU (bit-exact MT19937 in seq-gen's order); "host": by the
# reference's seq-gen on the host.  Both must reproduce the reference's output byte for byte.
@pytest.mark.parametrize("evolve", ["gpu", "host"])
@pytest.mark.parametrize("case", CASES)
def test_cli_output_identical_to_reference(case, evolve, tmp_path):
    if not (os.path.exists(CLI) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/RNAcode_cuda_det not built (needs /root/reference at build time)")
    parts = case.split(" ")
    fname, opts = parts[0], [p for p in parts[1:] if p]
    env = dict(os.environ, RNACODE_SEED="1")
    if evolve == "host":
        env["RNACODE_CUDA_EVOLVE"] = "host"
    res = subprocess.run([CLI, *opts, _input(fname, tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr
    assert _norm(res.stdout) == _norm(op.golden("cli_outputs")[case])


@pytest.mark.parametrize("workers", ["1", "5"])
@pytest.mark.parametrize("case", CASES)
def test_batched_pipeline_output_identical_to_reference(case, workers, tmp_path):
    """integration/rnacode_pipeline.c: all blocks of the file in one GPU batch, PhyML in forked workers, null
    alignments drawn on the GPU -- and still the reference's output, byte for byte."""
    if not (os.path.exists(PIPELINE) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/RNAcode_b200_det not built (needs /root/reference at build time)")
    parts = case.split(" ")
    fname, opts = parts[0], [p for p in parts[1:] if p]
    env = dict(os.environ, RNACODE_SEED="1", RNACODE_CUDA_WORKERS=workers)
    if workers == "5":
        env["RNACODE_CUDA_WINDOW"] = "4"  # several windows per file
    res = subprocess.run([PIPELINE, *opts, _input(fname, tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr
    assert _norm(res.stdout) == _norm(op.golden("cli_outputs")[case])


@pytest.mark.parametrize("case", ["genomic-preprocessed.maf --tabular -n 20", "genomic.maf --gtf -s -p 0.1 -n 40"])
def test_batched_pipeline_sharded_over_gpus(case):
    """RNACODE_CUDA_GPUS: the blocks of a window are dealt out over the devices (no exchange between them); the
    output does not depend on the number of devices."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not (os.path.exists(PIPELINE) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/RNAcode_b200_det not built (needs /root/reference at build time)")
    parts = case.split(" ")
    fname, opts = parts[0], [p for p in parts[1:] if p]
    env = dict(os.environ, RNACODE_SEED="1", RNACODE_CUDA_GPUS="all")
    res = subprocess.run([PIPELINE, *opts, os.path.join(EXAMPLES, fname)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr
    assert _norm(res.stdout) == _norm(op.golden("cli_outputs")[case])


def test_dead_tree_worker_is_fatal():
    """A PhyML worker that dies (here: killed while it holds block 3) ends the run with an error and a non-zero status, as a
    failing treeML ends the reference -- no silently missing blocks (round-1 advice)."""
    if not (os.path.exists(PIPELINE) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/RNAcode_b200_det not built (needs /root/reference at build time)")
    env = dict(os.environ, RNACODE_SEED="1", RNACODE_CUDA_WORKERS="4", RNACODE_CUDA_TEST_KILL_WORKER_AT="3")
    res = subprocess.run([PIPELINE, "--tabular", "-n", "20", os.path.join(EXAMPLES, "genomic-preprocessed.maf")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode != 0
    assert "tree worker was killed by signal 9" in res.stderr
    assert res.stdout == ""  # the window was not reported


def test_windows_overlap_in_the_batched_pipeline(tmp_path):
    """The workers of window w+1 are forked before window w is collected: in the verbose stage log every window but the
    first has its workers started before the previous window is done, and the output is still the reference's."""
    if not (os.path.exists(PIPELINE) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/RNAcode_b200_det not built (needs /root/reference at build time)")
    case = "genomic-preprocessed.maf --tabular -n 20"
    env = dict(os.environ, RNACODE_SEED="1", RNACODE_CUDA_WORKERS="3", RNACODE_CUDA_WINDOW="6", RNACODE_CUDA_VERBOSE="1")
    res = subprocess.run([PIPELINE, "--tabular", "-n", "20", os.path.join(EXAMPLES, "genomic-preprocessed.maf")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr
    assert _norm(res.stdout) == _norm(op.golden("cli_outputs")[case])
    wins = re.findall(r"window (\d+) of .*?workers forked at ([0-9.]+) s.*?window done at ([0-9.]+) s", res.stderr)
    assert len(wins) >= 5
    for (k, forked, _), (_, _, prev_done) in zip(wins[1:], wins[:-1]):
        assert float(forked) < float(prev_done), (k, forked, prev_done)


def test_gpu_rng_mode_pvalues_within_sampling_error():
    """GPU-RNG mode (RNACODE_CUDA_EVOLVE=philox): the HSS and their scores do not depend on the generator, and the
    p-values from the Gumbel fit of 2000 Philox-drawn null alignments agree with those of 2000 MT19937-drawn ones
    within sampling error (the fit's mu / lambda are estimated from the sample)."""
    import math
    exe = os.path.join(REFDIR, "RNAcode_cuda_det")
    if not (os.path.exists(exe) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/RNAcode_cuda_det not built (needs /root/reference at build time)")
    outs = {}
    for mode in ("gpu", "philox"):
        env = dict(os.environ, RNACODE_SEED="3", RNACODE_CUDA_EVOLVE=mode)
        res = subprocess.run([exe, "--tabular", "-n", "2000", os.path.join(EXAMPLES, "coding.aln")], capture_output=True,
                             text=True, env=env, timeout=600)
        assert res.returncode == 0, res.stderr
        outs[mode] = [l.split("\t") for l in res.stdout.strip().splitlines()]
    assert len(outs["gpu"]) == len(outs["philox"]) > 0
    n = 2000.0
    for a, b in zip(outs["gpu"], outs["philox"]):
        assert a[:-1] == b[:-1]  # same segment, strand, frame, coordinates, score
        pa, pb = float(a[-1]), float(b[-1])
        if min(pa, pb) > 0.5:
            assert abs(pa - pb) < 0.05
        else:
            # Expected sampling error: p = 1 - exp(-exp(-t)), t = (s - mu) / beta, so ln p ~ -t for small p, and the maximum-
            # likelihood fit of a Gumbel law from n samples has var(mu) = 1.1087 beta^2 / n, var(beta) = 0.6079 beta^2 / n,
            # cov = 0.2570 beta^2 / n  =>  var(ln p) = (1.1087 + 0.6079 t^2 + 0.514 t) / n per fit, twice that for the
            # difference of two independent fits.  At p = 1e-9 and n = 2000 that is 0.23 decades (one sigma); four sigma allowed.
            la, lb = math.log10(max(pa, 1e-300)), math.log10(max(pb, 1e-300))
            t = math.log(10.0) * max(abs(la), abs(lb))
            sigma = math.log10(math.e) * math.sqrt(2.0 * (1.1087 + 0.6079 * t * t + 0.514 * t) / n)
            assert abs(la - lb) < 4.0 * sigma + 0.01, (pa, pb, sigma)


EPS_CASES = sorted(op.golden("eps_outputs").keys())


@pytest.mark.parametrize("binary", ["RNAcode_cuda_det", "RNAcode_b200_det"])
@pytest.mark.parametrize("case", EPS_CASES)
def test_eps_plots_identical_to_reference(case, binary, tmp_path):
    """--eps: the reference's colorAln / backtrack walk rows of Sk_native that rc_pair_rows computes on the GPU
    (__wrap_backtrack, integration/rnacode_cuda_host.h).  Every hss-<n>.eps must be byte-identical to the file the
    unmodified reference writes (SHA-256 in tests/golden/eps_outputs.json.gz), and so must stdout."""
    import hashlib
    exe = os.path.join(REFDIR, binary)
    if not (os.path.exists(exe) and os.path.isdir(EXAMPLES)):
        pytest.skip("oracle/_ref/%s not built (needs /root/reference at build time)" % binary)
    gold = op.golden("eps_outputs")[case]
    parts = case.split(" ")
    fname, opts = parts[0], [p for p in parts[1:] if p]
    d = os.path.join(str(tmp_path), "eps")
    env = dict(os.environ, RNACODE_SEED="1")
    res = subprocess.run([exe, "--eps", "--eps-dir", d, *opts, os.path.join(EXAMPLES, fname)], capture_output=True, text=True,
                         env=env, timeout=600)
    assert res.returncode == 0, res.stderr
    assert _norm(res.stdout) == _norm(gold["stdout"])
    files = {n: hashlib.sha256(open(os.path.join(d, n), "rb").read()).hexdigest() for n in sorted(os.listdir(d))}
    assert files == gold["files"]
