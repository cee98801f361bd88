"""The decimal rendering of the device formatter (genometools_smax_b200/csrc/smax_dec.h)
is host+device code: compile it with gcc and check it against printf("%lu") at every
power of ten and of two and on millions of random values (no GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decimal_digits_and_rendering_match_printf(tmp_path):
    exe = str(tmp_path / "dec_check")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-Wall", "-Werror",
                    "-I", os.path.join(ROOT, "genometools_smax_b200", "csrc"),
                    os.path.join(ROOT, "tests", "dec_check.c"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.strip() == "dec ok"
