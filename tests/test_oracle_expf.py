"""CPU: the glibc-expf restatement (oracle/expf_glibc.h; the CUDA path carries a device copy of the
same algorithm) against this box's libm on a dense sample.  The exhaustive 2^32 sweep is
oracle/expf_sweep.c (result recorded in DESIGN.md)."""
import os
import subprocess

from helpers import ROOT


def test_expf_restatement_matches_libm_on_sample(tmp_path):
    exe = tmp_path / "expf_sweep"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "oracle", "expf_sweep.c"),
                    "-lm", "-o", str(exe)], check=True)
    # all negative floats from -2^-20 down to -128 (0xb5800000 .. 0xc3000000): the softmax domain
    out = subprocess.run([str(exe), "0xb5800000", "0xc3000000"], capture_output=True, text=True)
    assert "variant1(fma-reduce)=0" in out.stdout or "variant0(sse2)=0" in out.stdout, out.stdout
