#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/.  Run HERE (the build container), where
/root/reference exists; the outputs are committed because /root/reference does not travel to
the GPU box.

    python tests/golden/make_golden.py

Produces
  reference_fixtures/*.npy   the .npy files of the reference's own fixture generator
                             (scripts/gen_signal_fixtures.py, executed unmodified in a temp dir):
                             sosfilt_input.npy, resample_poly_output.npy, filtfilt_output.npy
  legacy_resample_taps.npy   the 31 literal taps of scir_signal::resample_poly
                             (crates/scir-signal/src/lib.rs:315-347), parsed as data
  scipy_vectors.npz          seeded inputs + SciPy outputs for lfilter(FIR) / upfirdn /
                             resample_poly / filtfilt(FIR) -- the API the reference lacks and
                             SciPy specifies
  provenance.json            versions used
"""
import json
import os
import re
import runpy
import shutil
import tempfile

import numpy as np
import scipy
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def reference_fixtures():
    out = os.path.join(HERE, "reference_fixtures")
    os.makedirs(out, exist_ok=True)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            runpy.run_path(os.path.join(REF, "scripts", "gen_signal_fixtures.py"), run_name="__main__")
        finally:
            os.chdir(cwd)
        for name in ("sosfilt_input.npy", "resample_poly_output.npy", "filtfilt_output.npy",
                     "butter_sos.npy"):
            shutil.copy(os.path.join(tmp, "fixtures", name), os.path.join(out, name))


def legacy_taps():
    src = open(os.path.join(REF, "crates/scir-signal/src/lib.rs")).read()
    m = re.search(r"const H: \[f64; 31\] = \[(.*?)\];", src, re.S)
    vals = [float(v.replace("_", "")) for v in re.findall(r"-?[0-9][0-9_.e+-]*", m.group(1))]
    assert len(vals) == 31, len(vals)
    np.save(os.path.join(HERE, "legacy_resample_taps.npy"), np.asarray(vals, dtype=np.float64))


def scipy_vectors():
    rng = np.random.RandomState(17)           # same seed as scipy's test_vs_convolve
    d = {}
    one32 = np.ones(1, np.float32)

    # lfilter, FIR branch, f32 in / f32 out (pass a as f32 so SciPy stays in f32)
    x = (rng.rand(3, 257).astype(np.float32) * 2 - 1)
    b = signal.firwin(63, 0.25).astype(np.float32)
    d["lfilter_x"], d["lfilter_b"] = x, b
    d["lfilter_y"] = signal.lfilter(b, one32, x, axis=-1)
    d["lfilter_y64"] = signal.lfilter(b.astype(np.float64), [1.0], x.astype(np.float64), axis=-1)
    zi = (rng.rand(3, 62).astype(np.float32) - 0.5)
    y, zf = signal.lfilter(b.astype(np.float64), [1.0], x.astype(np.float64), axis=-1,
                           zi=zi.astype(np.float64))
    d["lfilter_zi"], d["lfilter_y_zi"], d["lfilter_zf"] = zi, y, zf

    # upfirdn: a sweep of (len_h, len_x, up, down)
    cases = [(31, 200, 1, 1), (31, 200, 1, 2), (31, 200, 3, 2), (31, 200, 2, 3), (97, 333, 3, 2),
             (7, 50, 5, 3), (64, 129, 4, 7), (5, 3, 2, 1), (1, 17, 3, 4), (40, 10, 7, 5)]
    d["upfirdn_cases"] = np.asarray(cases, dtype=np.int64)
    for i, (lh, lx, up, down) in enumerate(cases):
        h = rng.randn(lh).astype(np.float32)
        xx = rng.randn(2, lx).astype(np.float32)
        d[f"upfirdn_{i}_h"], d[f"upfirdn_{i}_x"] = h, xx
        d[f"upfirdn_{i}_y64"] = signal.upfirdn(h.astype(np.float64), xx.astype(np.float64), up, down,
                                               axis=-1)

    # scipy test_vs_convolve (test_upfirdn.py:171-201): firwin(31, 1/down, hamming), randn(10000)
    xr = np.random.RandomState(17).randn(10000)
    d["vs_convolve_x"] = xr.astype(np.float32)
    for down, want_len in ((2, 5015), (11, 912), (79, 127)):
        h = signal.firwin(31, 1.0 / down, window="hamming")
        y = signal.upfirdn(h.astype(np.float32).astype(np.float64),
                           xr.astype(np.float32).astype(np.float64), 1, down)
        assert y.shape == (want_len,)
        d[f"vs_convolve_{down}_h"] = h.astype(np.float32)
        d[f"vs_convolve_{down}_y64"] = y

    # resample_poly with an explicit f32 window (BASELINE config 4 design, small)
    rp_cases = [(3, 2, 96, 500), (2, 3, 31, 32), (5, 7, 61, 211), (4, 2, 41, 100), (1, 3, 25, 77),
                (7, 1, 29, 41), (160, 147, 301, 300)]
    d["resample_cases"] = np.asarray(rp_cases, dtype=np.int64)
    for i, (up, down, lh, n) in enumerate(rp_cases):
        g = np.gcd(up, down)
        h = signal.firwin(lh, 1.0 / max(up // g, down // g), window=("kaiser", 5.0)).astype(np.float32)
        xx = (rng.rand(2, n).astype(np.float32) * 2 - 1)
        d[f"resample_{i}_h"], d[f"resample_{i}_x"] = h, xx
        d[f"resample_{i}_y32"] = signal.resample_poly(xx, up, down, axis=-1, window=h.copy())
        d[f"resample_{i}_y64"] = signal.resample_poly(xx.astype(np.float64), up, down, axis=-1,
                                                      window=h.astype(np.float64))

    # filtfilt with an FIR numerator: default odd padding, even, constant, none, custom padlen
    b = signal.firwin(31, 0.2).astype(np.float32)
    xx = (rng.rand(3, 400).astype(np.float32) * 2 - 1)
    d["filtfilt_b"], d["filtfilt_x"] = b, xx
    b64, x64 = b.astype(np.float64), xx.astype(np.float64)
    d["filtfilt_odd"] = signal.filtfilt(b64, [1.0], x64, axis=-1)
    d["filtfilt_even"] = signal.filtfilt(b64, [1.0], x64, axis=-1, padtype="even")
    d["filtfilt_const"] = signal.filtfilt(b64, [1.0], x64, axis=-1, padtype="constant")
    d["filtfilt_none"] = signal.filtfilt(b64, [1.0], x64, axis=-1, padtype=None)
    d["filtfilt_odd_pad10"] = signal.filtfilt(b64, [1.0], x64, axis=-1, padlen=10)
    # reference-structure (zero-state forward, reverse, zero-state forward, reverse),
    # the FIR analogue of scripts/gen_signal_fixtures.py:28
    d["filtfilt_refstyle"] = signal.lfilter(b64, [1.0], signal.lfilter(b64, [1.0], x64, axis=-1)[:, ::-1],
                                            axis=-1)[:, ::-1]

    np.savez_compressed(os.path.join(HERE, "scipy_vectors.npz"), **d)


def main():
    reference_fixtures()
    legacy_taps()
    scipy_vectors()
    sub = {}
    if os.path.exists(os.path.join(REF, ".SUBMODULES.json")):
        meta = json.load(open(os.path.join(REF, ".SUBMODULES.json")))
        sub = {"reference_commit": meta.get("commit"),
               "scipy_submodule_commit": next((m["commit"] for m in meta.get("submodules", [])
                                               if m.get("path") == "scipy"), None)}
    json.dump({"scipy": scipy.__version__, "numpy": np.__version__,
               "reference_requirements": [l for l in open(os.path.join(REF, "requirements.txt")).read().split()
                                          if l.lower().startswith(("numpy", "scipy"))],
               **sub},
              open(os.path.join(HERE, "provenance.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
