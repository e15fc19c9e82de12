#!/usr/bin/env python
"""Regenerates tests/golden/*.json.gz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

1. builds oracle/_ref/{RNAcode_ref,RNAcode_det,ref_probe} from /root/reference (oracle/Makefile target ref),
2. runs ref_probe (deterministic seeds via oracle/ref_wrap.c) on the reference's example alignments and
   on synthetic MAF written by rnacode_b200.synth (gappy, with N / lower-case / odd symbols, non-default
   --pars), and stores the JSON dumps gzip-compressed,
3. runs the deterministic reference CLI (RNAcode_det) on the examples with several option sets and stores
   its stdout (tests/golden/cli_outputs.json.gz) for the drop-in CLI comparison; with RC_GOLDEN_FULL=1 also BASELINE
   config 2 at its full -n 1000 (see long_cases).
The fixtures pin oracle/rnacode_oracle.c (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path.
"""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from rnacode_b200 import synth  # noqa: E402

REF = os.environ.get("RNACODE_REF", "/root/reference")
ORC = os.path.join(ROOT, "oracle")
PROBE = os.path.join(ORC, "_ref", "ref_probe")
DET = os.path.join(ORC, "_ref", "RNAcode_det")
TMP = os.path.join(ORC, "_ref", "tmp")


def probe(path, n, dump, extra=(), seed=1):
    env = dict(os.environ, RNACODE_SEED=str(seed))
    out = subprocess.run([PROBE, "-n", str(n), "--dump-samples", str(dump), *extra, path], check=True,
                         capture_output=True, env=env).stdout
    return json.loads(out)


def save(name, doc):
    p = os.path.join(HERE, name + ".json.gz")
    with gzip.GzipFile(p, "wb", mtime=0) as fh:
        fh.write(json.dumps(doc, separators=(",", ":")).encode())
    print("wrote", p, os.path.getsize(p) // 1024, "KiB")


def decorate(rows, rng):
    """Sprinkle N, lower-case and IUPAC symbols into a synthetic block (exercises ntMap / 'N' rules)."""
    rows = rows.copy()
    N, cols = rows.shape
    for sym, cnt in ((ord("N"), 6), (ord("R"), 3), (ord("Y"), 2), (ord("X"), 3)):
        for _ in range(cnt):
            s, c = rng.integers(0, N), rng.integers(0, cols)
            if rows[s, c] != synth.GAP:
                rows[s, c] = sym
    # one run of three X in a species row aligned to a reference codon, one NNN run in the reference
    rows[1, 30:33] = ord("X")
    rows[0, 60:63] = ord("N")
    return rows


def mixed_blocks():
    """36 synthetic blocks of mixed shape (3-33 rows, 30-900 columns, with and without gaps): at -n 30 they exercise every
    DP route of the library (sample-major resident / streamed / chunked, row-major registers / chain) end to end."""
    rng = np.random.default_rng(424242)
    shapes = []
    for _ in range(36):
        N = int(rng.choice([3, 4, 5, 8, 10, 12, 17, 20, 26, 33]))
        cols = int(rng.choice([30, 45, 60, 90, 120, 200, 330, 450, 600, 900]))
        if N >= 20 and cols > 450:
            cols = 450
        shapes.append((N, cols))
    return [synth.synth_block(77, i, N, cols, gap_rate=float(rng.choice([0.0, 0.0067, 0.02]))) for i, (N, cols) in enumerate(shapes)]


EPS_CASES = (("coding.aln", ["--tabular"]), ("coding.maf", ["--gtf"]),
             ("genomic-preprocessed.maf", ["--tabular", "-n", "20", "--eps-cutoff", "0.9"]))


def dupnames_maf(ex, path):
    """examples/coding.maf with its second row renamed to the first row's name: the reference still scores such a block
    (sortAln's swaps, src/misc.c:150-171, leave an arrangement of the simulated rows), and so must the drop-ins."""
    lines = open(os.path.join(ex, "coding.maf")).read().split("\n")
    srows = [i for i, ln in enumerate(lines) if ln.startswith("s ")]
    first = lines[srows[0]].split()[1]
    second = lines[srows[1]].split()[1]
    lines[srows[1]] = lines[srows[1]].replace(second, first.ljust(len(second)), 1) if len(first) <= len(second) else \
        lines[srows[1]].replace(second, first, 1)
    with open(path, "w") as fh:
        fh.write("\n".join(lines))
    return path


def eps_golden(ex):
    """--eps plots of the unmodified reference (deterministic seeds): stdout and the SHA-256 of every hss-<n>.eps."""
    import hashlib
    import shutil
    doc = {}
    for f, opts in EPS_CASES:
        d = os.path.join(TMP, "eps_ref")
        shutil.rmtree(d, ignore_errors=True)
        env = dict(os.environ, RNACODE_SEED="1")
        out = subprocess.run([DET, "--eps", "--eps-dir", d, *opts, os.path.join(ex, f)], check=True, capture_output=True,
                             env=env).stdout.decode()
        files = {n: hashlib.sha256(open(os.path.join(d, n), "rb").read()).hexdigest() for n in sorted(os.listdir(d))}
        doc[f + " " + " ".join(opts)] = {"stdout": out, "files": files}
    return doc


LONG_CLI_CASE = "genomic.maf --gtf --best-only -n 1000"  # BASELINE.json configs[1], the headline command


def load(name):
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "rb") as fh:
        return json.loads(fh.read())


def long_cases(ex, cli):
    """BASELINE config 2 at its full -n 1000 (14 minutes of one core each, the 10 x 4806 block dominates): the reference's
    stdout for the headline command and, through ref_probe, the 1000 per-sample maxima, the Gumbel fit and the p-values of
    every block.  Regenerated only with RC_GOLDEN_FULL=1; otherwise the committed entries are kept."""
    if os.environ.get("RC_GOLDEN_FULL") == "1":
        env = dict(os.environ, RNACODE_SEED="1")
        opts = LONG_CLI_CASE.split(" ")[1:]
        cli[LONG_CLI_CASE] = subprocess.run([DET, *opts, os.path.join(ex, "genomic.maf")], check=True, capture_output=True,
                                            env=env).stdout.decode()
        doc = probe(os.path.join(ex, "genomic.maf"), 1000, 0)
        for blk in doc["blocks"]:
            blk.pop("samples", None)
        save("genomic_maf_n1000", doc)
    else:
        cli[LONG_CLI_CASE] = load("cli_outputs")[LONG_CLI_CASE]
        print("kept the committed -n 1000 cases (set RC_GOLDEN_FULL=1 to regenerate them: 2 x 14 core-minutes)")


def main():
    subprocess.run(["make", "-C", ORC, "-j8", "ref"], check=True, stdout=subprocess.DEVNULL)
    os.makedirs(TMP, exist_ok=True)
    ex = os.path.join(REF, "examples")
    save("coding_aln", probe(os.path.join(ex, "coding.aln"), 100, 100))
    save("noncoding_aln", probe(os.path.join(ex, "noncoding.aln"), 50, 50))
    save("coding_maf", probe(os.path.join(ex, "coding.maf"), 20, 20))
    save("noncoding_maf", probe(os.path.join(ex, "noncoding.maf"), 20, 20))
    save("genomic_maf", probe(os.path.join(ex, "genomic.maf"), 3, 3))
    save("genomic_pre_maf", probe(os.path.join(ex, "genomic-preprocessed.maf"), 8, 8))
    save("coding_aln_pars", probe(os.path.join(ex, "coding.aln"), 20, 20, extra=("--pars", "-9.5,-3.25,-1.5,-50")))

    rng = np.random.default_rng(7)
    blocks = [synth.synth_block(11, i, N, cols, gap_rate=gr)
              for i, (N, cols, gr) in enumerate([(10, 120, 0.0067), (10, 120, 0.03), (5, 200, 0.02), (12, 90, 0.05),
                                                 (3, 150, 0.01), (10, 333, 0.02), (25, 100, 0.02), (30, 75, 0.03)])]
    blocks[1] = decorate(blocks[1], rng)
    blocks[3] = decorate(blocks[3], rng)
    p = os.path.join(TMP, "synth_gappy.maf")
    synth.to_maf(blocks, p)
    save("synth_gappy", probe(p, 12, 12))
    blocks = [synth.synth_block(12, i, 10, 120, gap_rate=0.0) for i in range(3)]
    p = os.path.join(TMP, "synth_gapfree.maf")
    synth.to_maf(blocks, p)
    save("synth_gapfree", probe(p, 12, 12))

    # rows of the reference's pairwise matrices Sk_native / Sk_native_rev (what backtrack() walks for --eps)
    sk = {"coding_aln": probe(os.path.join(ex, "coding.aln"), 1, 0, extra=("--sk-rows", "8")),
          "synth_gappy": probe(os.path.join(TMP, "synth_gappy.maf"), 1, 0, extra=("--sk-rows", "6"))}
    save("sk_rows", sk)

    # reference CLI outputs (deterministic seeds) for the drop-in comparison
    cli = {}
    for f, opts in (("coding.aln", []), ("coding.aln", ["--tabular"]), ("coding.aln", ["--gtf"]),
                    ("coding.maf", ["--gtf"]), ("coding.maf", ["--tabular", "-b"]), ("noncoding.aln", ["--tabular"]),
                    ("genomic-preprocessed.maf", ["--tabular", "-n", "20"]),
                    ("genomic-preprocessed.maf", ["--gtf", "-r", "-n", "20"]),
                    ("genomic-preprocessed.maf", ["--tabular", "-s", "-p", "0.05", "-n", "20"]),
                    # more than 32 samples under --stop-early: the batched pipeline samples in two rounds
                    ("genomic-preprocessed.maf", ["--tabular", "-s", "-p", "0.05", "-n", "70"]),
                    ("genomic.maf", ["--gtf", "-s", "-p", "0.1", "-n", "40"]),
                    ("coding.aln", ["--tabular", "-c", "-9.5,-3.25,-1.5,-50"])):
        env = dict(os.environ, RNACODE_SEED="1")
        out = subprocess.run([DET, *opts, os.path.join(ex, f)], check=True, capture_output=True, env=env).stdout.decode()
        cli[f + " " + " ".join(opts)] = out
    # synthetic MAF of mixed shapes through the unmodified reference (the tests rebuild the same file with mixed_blocks())
    p = os.path.join(TMP, "synth_mixed.maf")
    synth.to_maf(mixed_blocks(), p)
    env = dict(os.environ, RNACODE_SEED="1")
    cli["synthetic:mixed --tabular -n 30"] = subprocess.run([DET, "--tabular", "-n", "30", p], check=True, capture_output=True,
                                                            env=env).stdout.decode()
    p = dupnames_maf(ex, os.path.join(TMP, "dupnames.maf"))
    cli["synthetic:dupnames --tabular -n 20"] = subprocess.run([DET, "--tabular", "-n", "20", p], check=True, capture_output=True,
                                                               env=env).stdout.decode()
    long_cases(ex, cli)
    save("cli_outputs", cli)
    save("eps_outputs", eps_golden(ex))


if __name__ == "__main__":
    main()
