"""Builds scir_b200/lib/libscir_b200.so in-tree with nvcc for sm_100a (make -C scir_b200/csrc).

The reference has no build step for its GPU code (a PTX string JIT-compiled at run time,
crates/scir-gpu/src/lib.rs:814-824); a replacement crate's build.rs would run the same nvcc lines
(see INTEGRATION.md).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libscir_b200.so")


def build(verbose: bool = False, jobs: int = 8) -> str:
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libscir_b200.so failed (see output above)")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
