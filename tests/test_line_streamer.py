"""CPU fuzz of the gather's line buffer (versalignlib_b200/csrc/va_line_streamer.h: non-temporal stores of whole lines,
plain copies at the ends, a direct path for long sequences) against memcpy, with guard bytes around the destination.
The class is host-only C++, so the test builds a small harness with the host compiler -- once as it ships (SSE2) and once
with the portable fallback."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "line_streamer_fuzz.cpp")
INC = os.path.join(ROOT, "versalignlib_b200", "csrc")


@pytest.mark.parametrize("flags", [[], ["-mno-sse2", "-mno-sse", "-mfpmath=387"]], ids=["sse2", "portable"])
def test_line_streamer_matches_memcpy(tmp_path, flags):
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "fuzz")
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", *flags, "-I", INC, SRC, "-o", exe], capture_output=True, text=True)
    if r.returncode != 0 and flags:
        pytest.skip("this compiler cannot build without SSE: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe, "1500"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
