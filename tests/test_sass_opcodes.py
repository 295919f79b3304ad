"""The Blackwell-native claim, checked mechanically on the built library (no GPU needed): the tensor kernel carries
tcgen05 MMAs with TMEM loads/stores, the streaming kernels carry TMA bulk copies and packed FFMA2, the f64 twin carries
DFMA, and the only cubin architecture is sm_100a.  tools/sass_summary.py writes the same table to
profiles/r02_sass_opcodes.txt (committed)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def families():
    import sass_summary
    from scir_b200 import _lib as L
    from scir_b200 import build
    build.build()
    counts, short, arch = sass_summary.summarise(L.LIB_PATH)
    return sass_summary.family(counts, short), arch


def test_only_sm100a(families):
    _, arch = families
    assert arch == {"sm_100a"}


def test_tensor_kernel_is_tcgen05(families):
    fam, _ = families
    t = fam["fir_toeplitz_kernel"]
    assert t["UTCHMMA"] > 0 and t["UTCBAR"] > 0          # tcgen05.mma + tcgen05.commit
    assert t["LDTM"] > 0 and t["STTM"] > 0               # accumulators read from / operands written to TMEM
    assert t["UBLKCP"] > 0 and t["SYNCS"] > 0            # TMA bulk copies on mbarriers
    assert t["FFMA"] == 0 and t["FFMA2"] == 0            # no CUDA-core contraction hiding in the tensor kernel


def test_streaming_kernels_use_tma_and_ffma2(families):
    fam, _ = families
    for name in ("fir_tile_kernel", "upfirdn_stream_kernel", "upfirdn_tile_kernel"):
        k = fam[name]
        assert k["UBLKCP"] > 0 and k["SYNCS"] > 0, name
        assert k["FFMA2"] > 0, name
    assert fam["fir_f64_kernel"]["DFMA"] > 0
    for name, k in fam.items():
        if name != "fir_toeplitz_kernel":
            assert k["UTCHMMA"] == 0, name


def test_committed_summary_is_current(families):
    """profiles/r02_sass_opcodes.txt must be regenerated when the kernels change (python tools/sass_summary.py)."""
    fam, _ = families
    path = os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
    assert os.path.exists(path), "run python tools/sass_summary.py"
    txt = open(path).read()
    row = [ln for ln in txt.splitlines() if ln.startswith("fir_toeplitz_kernel ")]
    assert row, "fir_toeplitz_kernel row missing"
    assert int(row[0].split()[2]) == fam["fir_toeplitz_kernel"]["UTCHMMA"]
