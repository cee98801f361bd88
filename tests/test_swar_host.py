"""The bit-parallel core of K1/K2 (genometools_smax_b200/csrc/smax_swar.h) is
host+device code: compile it with gcc and check it byte for byte against a
scalar restatement on millions of random / adversarial chunks (no GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swar_chunk_functions_match_scalar(tmp_path):
    exe = str(tmp_path / "swar_check")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-Wall", "-Werror",
                    "-I", os.path.join(ROOT, "genometools_smax_b200", "csrc"),
                    os.path.join(ROOT, "tests", "swar_check.c"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.strip() == "swar ok"
