"""x87 extended-precision emulation (ngsld_b200/csrc/fp80.cuh, the arithmetic behind the bit-exact r2_ExpG column)
against the host FPU's native long double: compiled as plain C++ and run on the CPU."""
import os
import subprocess

import helpers as H


def test_fp80_emulation_matches_native_long_double(tmp_path):
    exe = str(tmp_path / "fp80_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(H.HERE, "native", "fp80_check.cpp")])
    out = subprocess.run([exe, "400000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "0 mismatches" in out.stdout
