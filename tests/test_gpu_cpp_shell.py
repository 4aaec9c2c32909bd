"""-m gpu: builds tests/cpp/test_host_shell.cpp (the reference's gtest cases written against
the C++ host shell, fastdem_b200/host/fastdem/*.hpp) with g++ against libfastdem_b200.so and
runs it on the device."""
import subprocess
from pathlib import Path

import pytest

from fastdem_b200 import capi

REPO = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_cpp_host_shell(tmp_path, fdem):
    exe = tmp_path / "test_host_shell"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(REPO / "include"), "-I", str(REPO / "fastdem_b200" / "host"),
           str(REPO / "tests" / "cpp" / "test_host_shell.cpp"), "-o", str(exe),
           str(capi.LIB_PATH), f"-Wl,-rpath,{capi.LIB_PATH.parent}"]
    subprocess.run(cmd, check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL C++ SHELL TESTS PASSED" in r.stdout
