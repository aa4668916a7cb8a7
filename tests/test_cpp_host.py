"""The C++ host layer (include/hashdag_b200.hpp) compiles and links against the C-ABI library; on a GPU it runs the
resurrected intents of the reference's test/test.cpp through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api.bin")
LIBDIR = os.path.join(ROOT, "vkhashdag_b200", "csrc")


def build():
    import vkhashdag_b200 as v
    v.lib()
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                           "-L", LIBDIR, "-lhashdag_b200", f"-Wl,-rpath,{LIBDIR}"])


def test_cpp_host_layer_compiles_and_links():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_layer_runs_reference_test_intents():
    build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp host api: OK" in out.stdout
