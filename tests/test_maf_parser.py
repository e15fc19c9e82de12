"""(f3) The memory-mapped MAF reader of the batched driver (integration/rnacode_maf_mmap.h) against the reference's
read_maf (src/rnaz_utils.c:132-234): oracle/_ref/maf_parse_check parses a file with both and compares every block field by
field (names, upper-cased sequences, start, length, source length, strand)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
CHECK = os.path.join(REFDIR, "maf_parse_check")

pytestmark = pytest.mark.skipif(not os.path.exists(CHECK), reason="oracle/_ref/maf_parse_check not built (needs /root/reference)")


def _check(path):
    res = subprocess.run([CHECK, path], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    tag, blocks, rows = res.stdout.split()
    assert tag == "OK"
    return int(blocks), int(rows)


@pytest.mark.parametrize("name,blocks", [("coding.maf", 1), ("noncoding.maf", 1), ("genomic.maf", 11),
                                         ("genomic-preprocessed.maf", 34)])
def test_examples(name, blocks):
    assert _check(os.path.join(REFDIR, "examples", name))[0] == blocks


def test_synthetic_mixed(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    from rnacode_b200 import synth
    p = os.path.join(str(tmp_path), "mixed.maf")
    synth.to_maf(make_golden.mixed_blocks(), p)
    assert _check(p)[0] == 36


def test_irregular_lines(tmp_path):
    """Comments, blank lines, i / e / q lines, unknown lines, tabs, CR LF, lower case, no newline at the end."""
    p = os.path.join(str(tmp_path), "nasty.maf")
    with open(p, "w") as fh:
        fh.write("##maf version=1\n# c\n\na score=1\ns a.chr1\t10 6 + 100 acgt-NN\r\ns b.chr2 0 7 - 50 ACGTTnn\n"
                 "i b.chr2 N 0 C 0\n\nq b.chr2 99999\n#x\na score=2\ns a 1 3 + 9 ACG\ns b 2 3 + 9 acg\ne c 0 0 + 0 I\n"
                 "z junk line\ns c 3 3 - 9 A-G")
    assert _check(p) == (2, 5)


@pytest.mark.parametrize("body", [
    "a score=1\ns a 0 3 + 9 ACG extra\ns b 0 3 + 9 ACG\n",       # 8 fields
    "a score=1\ns a 0 3 + 9\n",                                   # 6 fields
    "a score=1\ns a x 3 + 9 ACG\ns b 0 3 + 9 ACG\n",             # start is not an integer
    "a score=1\ns a 0 y + 9 ACG\n",                               # length is not an integer
    "a score=1\ns a 0 3 + z ACG\n",                               # source length is not an integer
    "a score=1\ns a 0 3 ? 9 ACG\n",                               # strand
    "a score=1\ns a 0 3 + 9 ACG\ns b 0 4 + 9 ACGT\n",            # unequal lengths
], ids=["8fields", "6fields", "start", "length", "srcsize", "strand", "unequal"])
def test_malformed_input_fails_like_the_reference(body, tmp_path):
    """Both readers stop with the same message and a non-zero exit status."""
    p = os.path.join(str(tmp_path), "bad.maf")
    with open(p, "w") as fh:
        fh.write("##maf version=1\n" + body)
    ref = subprocess.run([CHECK, "--reference-only", p], capture_output=True, text=True, timeout=60)
    got = subprocess.run([CHECK, "--mapped-only", p], capture_output=True, text=True, timeout=60)
    assert ref.returncode != 0 and got.returncode != 0
    assert got.stderr == ref.stderr and ref.stderr.strip() != ""


def test_block_without_sequences_does_not_end_the_file(tmp_path):
    """Two consecutive 'a' lines: the reference's read_maf dereferences the empty block and crashes; the mapped reader skips
    it and goes on with the blocks that follow instead of treating it as the end of the input."""
    p = os.path.join(str(tmp_path), "empty_block.maf")
    with open(p, "w") as fh:
        fh.write("##maf version=1\na score=0\na score=1\ns a.c 0 6 + 100 ACGTAC\ns b.c 0 6 + 100 ACGTAC\ns c.c 0 6 + 100 ACGTAA\n\n"
                 "a score=2\n\na score=3\ns a.c 0 3 + 100 ACG\ns b.c 0 3 + 100 ACG\n")
    got = subprocess.run([CHECK, "--mapped-only", p], capture_output=True, text=True, timeout=60)
    assert got.returncode == 0 and got.stdout.split() == ["OK", "2", "5"], got.stdout + got.stderr
