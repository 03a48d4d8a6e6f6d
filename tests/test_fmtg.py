"""csrc/xm_fmtg.h (a 32-bit float as "%g" prints it, used for BAM float aux values) against the C library, on the CPU:
every 4099th of the 2^32 bit patterns plus the neighbourhoods where the notation changes.  scripts/check_fmtg_all.sh runs all
2^32 (about five minutes on eight cores; last run: 0 differ)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fmt_g_equals_printf_on_a_stride_of_all_floats(tmp_path):
    exe = str(tmp_path / "fmtg_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "emu", "fmtg_check.cpp")])
    for stride in ("4099", "65521"):
        r = subprocess.run([exe, stride, "4"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert " 0 differ" in r.stdout
