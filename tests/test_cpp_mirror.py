"""Builds and runs tests/cpp/mirror_test.cpp against include/scir_b200.hpp + libscir_b200.so: the C++
host-side mirror of the reference's Rust surface (the reference is compiled code; its toolchain is
absent here).  CPU run: Device::Cuda fails loudly, integer plans work.  GPU run: known-answer vectors."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "scir_b200", "lib")
EXE = os.path.join(ROOT, "build", "mirror_test")


def _build():
    from scir_b200 import build
    build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"), "-o", EXE,
           "-L", LIBDIR, "-lscir_b200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_mirror_fails_loudly_without_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present; the loud-failure leg is for the CPU-only container")
    _build()
    res = subprocess.run([EXE, "cpu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "backend not available" in res.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu():
    _build()
    res = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr


def test_copy_pool_host_staging_logic():
    """The pinned-ring staging of the *_host entry points is plain host code (scir_b200/csrc/copy_pool.hpp): streaming copies
    at every alignment and strided 2-D submissions over a thread pool, tested here without a GPU."""
    exe = os.path.join(ROOT, "build", "copy_pool_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "copy_pool_test.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", src, "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
