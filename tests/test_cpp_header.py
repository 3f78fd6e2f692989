"""include/fmx.hpp type-checks against include/fmx.h (no GPU, no link): the compiled facade is exercised on a GPU by
tests/test_gpu_parity.py::test_cpp_facade; this keeps the header's whole surface -- the fused batched query and the
multi-GPU group included -- compiling on every CPU run."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not installed")
@pytest.mark.parametrize("src", ["facade_surface.cpp", "test_api.cpp"])
def test_cpp_facade_type_checks(src):
    res = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "tests", "cpp", src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_c_header_is_plain_c():
    """the boundary is a C ABI: fmx.h must compile as C (gcc -std=c99), not only as C++"""
    if shutil.which("gcc") is None:
        pytest.skip("gcc not installed")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "fmx.h")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
