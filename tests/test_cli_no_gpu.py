"""The product has no CPU path: on a machine without a CUDA device both reference-side bindings must stop with a
clear message instead of computing anything on the host."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.parametrize("binary", ["RNAcode_cuda", "RNAcode_b200"])
def test_cli_fails_loudly_without_a_gpu(binary):
    exe = os.path.join(REFDIR, binary)
    aln = os.path.join(REFDIR, "examples", "coding.aln")
    if not (os.path.exists(exe) and os.path.exists(aln)):
        pytest.skip("oracle/_ref binaries not built (need /root/reference at build time)")
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    res = subprocess.run([exe, "--tabular", "-n", "5", aln], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "no usable CUDA device" in res.stderr
    assert res.stdout.strip() == ""  # nothing was scored
