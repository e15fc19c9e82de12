"""k_dp_regtu brings the deferred state U up to date with a closed form instead of n additions of omega.
tools/catch_up_check.c holds the same statements in C and compares them with the step-by-step loop, bit for bit, on random
(value, omega = -2^k, n): zero, tiny, huge, positive and negative values, runs that cross many binades."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_closed_form_equals_the_loop(tmp_path):
    exe = os.path.join(str(tmp_path), "catch_up_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "catch_up_check.c"), "-lm"], check=True)
    res = subprocess.run([exe, "1500000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "mismatches 0" in res.stdout


def test_cuda_source_holds_the_same_closed_form():
    """The C check is only worth something while the kernel uses the same statements."""
    src = open(os.path.join(ROOT, "rnacode_b200", "csrc", "rc_kernels.cuh")).read()
    body = src[src.index("__device__ __forceinline__ float catch_up("):]
    body = body[:body.index("\n}\n")]
    for stmt in ("const int shift = kexp + 150 - expf;", "if (expf == 0 || shift < 0)", "(bits & 0x7f800000u) + 0x00800000u",
                 "__fmaf_rn(w, (float)n, u)", "if (fabsf(r) <= B) return r;", "(int)((bits & 0x7fffffu) | 0x800000u)",
                 "((bits >> 31) ? -m : m) + (1 << 24)", "shift >= 25 ? 0 : (num >> shift)", "__fmaf_rn(w, (float)(J + 1), u)",
                 "n -= J + 1;"):
        assert stmt in body, stmt
