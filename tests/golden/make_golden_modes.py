#!/usr/bin/env python3
"""Golden vectors for the signal-extension modes of upfirdn / resample_poly, generated with the
importable SciPy (the reference's behavioural spec for the FIR routes it lacks: scripts/gen_signal_fixtures.py
uses scipy.signal the same way).  Run in this container:  python tests/golden/make_golden_modes.py
Writes tests/golden/scipy_modes.npz; provenance (versions) goes into the npz itself."""
import os

import numpy as np
import scipy
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = ["constant", "symmetric", "edge", "smooth", "wrap", "reflect", "antisymmetric", "antireflect", "line"]


def main():
    rng = np.random.RandomState(2024)
    d = {"scipy_version": np.asarray(scipy.__version__), "numpy_version": np.asarray(np.__version__),
         "modes": np.asarray(MODES)}
    # (len_h, len_x, up, down): interior-dominated, rational both ways, x shorter than the filter reach
    # (multiple reflections: _upfirdn_apply.pyx:121-128), plain FIR
    cases = [(31, 200, 3, 2), (13, 50, 2, 3), (40, 10, 7, 5), (97, 333, 1, 1), (33, 64, 4, 1), (21, 5, 1, 2)]
    d["cases"] = np.asarray(cases, dtype=np.int64)
    for i, (lh, lx, up, down) in enumerate(cases):
        h = rng.randn(lh).astype(np.float32)
        x = rng.randn(2, lx).astype(np.float32)
        d[f"u{i}_h"], d[f"u{i}_x"] = h, x
        for m in MODES:
            d[f"u{i}_{m}"] = signal.upfirdn(h.astype(np.float64), x.astype(np.float64), up, down, axis=-1, mode=m)
        d[f"u{i}_constant_cval"] = signal.upfirdn(h.astype(np.float64), x.astype(np.float64), up, down, axis=-1,
                                                  mode="constant", cval=0.75)
    # resample_poly padtypes (scipy/signal/_signaltools.py:3921-3957), f64 reference on f32-representable data
    rcases = [(3, 2, 96, 500), (2, 3, 31, 120), (5, 7, 61, 211)]
    d["rcases"] = np.asarray(rcases, dtype=np.int64)
    pads = MODES + ["mean", "median", "minimum", "maximum"]
    d["padtypes"] = np.asarray(pads)
    for i, (up, down, lh, n) in enumerate(rcases):
        h = signal.firwin(lh, 1.0 / max(up, down), window=("kaiser", 5.0)).astype(np.float32)
        x = (rng.rand(2, n).astype(np.float32) * 2 + 0.5)                  # non-zero mean: padtypes differ visibly
        d[f"r{i}_h"], d[f"r{i}_x"] = h, x
        for pt in pads:
            d[f"r{i}_{pt}"] = signal.resample_poly(x.astype(np.float64), up, down, axis=-1,
                                                   window=h.astype(np.float64), padtype=pt)
        d[f"r{i}_constant_cval"] = signal.resample_poly(x.astype(np.float64), up, down, axis=-1,
                                                        window=h.astype(np.float64), padtype="constant", cval=0.5)
    np.savez_compressed(os.path.join(HERE, "scipy_modes.npz"), **d)
    print("wrote scipy_modes.npz with", len(d), "arrays; scipy", scipy.__version__)


if __name__ == "__main__":
    main()
