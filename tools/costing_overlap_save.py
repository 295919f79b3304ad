#!/usr/bin/env python3
"""Costing of an overlap-save (block-FFT) path for long filters on the tensor cores -- VERDICT r1 task 6.

Question: config 3 (K = 4097) runs at the chip's deliverable MMA rate but executes 25 344 tensor flop per output; would
a block-FFT convolution whose DFT stages are tcgen05 GEMMs (128-point DFT matrices, two stages -> 16 384-point blocks,
the same block-scaled FP16 x 3 split as the Toeplitz kernel) reach <= 8 ms AT THE PATH'S TOLERANCE
(max|err| <= 1e-5 * sum|h| * max|x|)?

This script (numpy, CPU) answers both halves:
  1. flops: executed tensor flop per output for the FFT route, against the Toeplitz kernel's;
  2. error: an emulation of the four chained split-precision GEMM stages (forward 2, inverse 2) on the inputs the
     parity tests use -- uniform noise, a constant, a full-scale tone -- against an f64 convolution.
Output is committed as profiles/r02_overlap_save_costing.txt and summarised in DESIGN.md.
"""
import numpy as np

N1 = 128
N = N1 * N1                       # 16 384-point blocks
K = 4097
L = N - (K - 1)                   # valid outputs per block


def firwin(ntaps, cutoff):
    m = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = cutoff * np.sinc(cutoff * m) * np.hamming(ntaps)
    return (h / h.sum()).astype(np.float32).astype(np.float64)


def split16(a, terms, axis=None):
    """Block-scaled FP16 split of a real array: returns the value the tensor core effectively multiplies
    (sum of `terms` fp16 terms after scaling max|a| into [2^14, 2^15)).  axis=None: one scale per tile (what
    fir_toeplitz.cu does per slab); axis=0/1: one scale per column / row."""
    m = np.max(np.abs(a), axis=axis, keepdims=True)
    m = np.where(m == 0, 1.0, m)
    scale = 2.0 ** (14 - np.floor(np.log2(m)))
    r = a * scale
    out = np.zeros_like(r)
    for _ in range(terms):
        t = (r - out).astype(np.float16).astype(np.float64)
        out = out + t
    return out / scale


def mm_split(a, b, terms_kept, axis_b=None):
    """a @ b with both operands split into two fp16 terms each and only the hh, hm, mh products kept (terms_kept=3),
    or all four (4), or three terms per operand and six products (6)."""
    if terms_kept == 6:
        return split16(a, 3) @ split16(b, 3, axis_b)                  # hh hm mh mm hl lh ~ 3-term operands (upper bound on quality)
    ah, bh = split16(a, 1), split16(b, 1, axis_b)
    a2, b2 = split16(a, 2), split16(b, 2, axis_b)
    full = a2 @ b2
    if terms_kept == 4:
        return full
    return full - (a2 - ah) @ (b2 - bh)                               # drop the mm product


def dft_mats():
    k = np.arange(N1)
    w = np.exp(-2j * np.pi * np.outer(k, k) / N1)
    return w.real.copy(), w.imag.copy()


C, S = dft_mats()
TW = np.exp(-2j * np.pi * np.outer(np.arange(N1), np.arange(N1)) / N)       # twiddles W_N^(n2*k1)


def cmm(ar, ai, br, bi, terms, axis_b=None):
    """complex matmul (ar + i ai)(br + i bi) as four split real GEMMs with FP32-like (here f64) accumulation"""
    rr = mm_split(ar, br, terms, axis_b) - mm_split(ai, bi, terms, axis_b)
    ri = mm_split(ar, bi, terms, axis_b) + mm_split(ai, br, terms, axis_b)
    return rr, ri


def fft16k(z, terms, inverse=False, axis_b=None):
    """16 384-point DFT of complex z as F128 . X (.) twiddle . F128 with split GEMMs; f32 rounding of stage outputs."""
    x = z.reshape(N1, N1)                       # x[n1, n2], n = 128 n1 + n2
    s = -S if not inverse else S                # exp(-i..) = C - iS ... S holds the sign already: w.imag = -sin
    ci, si = C, (S if not inverse else -S)
    # stage 1: DFT over n1 (rows): Y[k1, n2] = sum_n1 W128^(n1 k1) x[n1, n2]
    yr, yi = cmm(ci, si, x.real, x.imag, terms, axis_b)
    y = (yr + 1j * yi).astype(np.complex64).astype(np.complex128)           # stage output leaves TMEM as f32
    y = y * (TW if not inverse else np.conj(TW))                             # twiddle, FP32 CUDA cores (exact enough)
    y = y.astype(np.complex64).astype(np.complex128)
    # stage 2: DFT over n2 (columns): Z[k1, k2] = sum_n2 y[k1, n2] W128^(n2 k2)
    zr, zi = cmm(y.real, y.imag, ci, si, terms, None if axis_b is None else 1 - axis_b)
    out = (zr + 1j * zi).astype(np.complex64).astype(np.complex128)
    return out.T.reshape(-1)                    # index k = k1 + 128 k2


def check_fft():
    rng = np.random.RandomState(0)
    z = rng.randn(N) + 1j * rng.randn(N)
    ref = np.fft.fft(z)
    got = fft16k(z, 6)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5, np.abs(got - ref).max() / np.abs(ref).max()


def overlap_save_block(xa, xb, hspec, terms, axis_b):
    """two real blocks packed as one complex signal (halves the GEMM count), filtered, unpacked"""
    z = xa + 1j * xb
    Z = fft16k(z, terms, False, axis_b)
    Y = (Z * hspec).astype(np.complex64).astype(np.complex128)
    y = np.conj(fft16k(np.conj(Y), terms, False, axis_b)) / N                # inverse via conj trick, same GEMMs
    return y.real[K - 1:], y.imag[K - 1:]


def main():
    check_fft()
    h = firwin(K, 0.01)
    hspec = np.fft.fft(np.r_[h, np.zeros(N - K)])
    rng = np.random.RandomState(1)
    t = np.arange(N)
    cases = {
        "uniform U[-1,1)": rng.rand(2, N) * 2 - 1,
        "constant 0.7": np.full((2, N), 0.7),
        "tone 0.9 sin(2 pi 0.003 n) (pass band)": np.tile(0.9 * np.sin(2 * np.pi * 0.003 * t), (2, 1)),
        "tone + noise 1e-3": np.tile(0.9 * np.sin(2 * np.pi * 0.003 * t), (2, 1)) + 1e-3 * rng.randn(2, N),
    }
    print(f"block {N} = {N1} x {N1}, K = {K}, valid outputs per block L = {L} ({L / N:.0%})")
    # ---- flops -------------------------------------------------------------------------------------------------
    gemm = 2.0 * N1 ** 3                                                    # one real 128^3 GEMM
    per_two_blocks = 16 * gemm                                              # fwd 2 complex stages + inv 2 complex stages, 4 real GEMMs each
    for terms in (3, 4, 6):
        ex = per_two_blocks * terms / (2 * L)
        pmax = (K - 1 + 127) // 128
        ksteps = sum(8 - (max(0, 128 * pb - (K - 1)) >> 4) for pb in range(pmax + 1))
        toe = 3 * ksteps * (2.0 * 128 * 128 * 16) / (128 * 128)
        outs = 256 * (1 << 22)
        for peak, name in ((1646e12, "cuBLAS burst"), (1370e12, "sustained")):
            print(f"  split x{terms}: {ex:8.0f} executed tensor flop/output (Toeplitz x3: {toe:.0f}) -> config 3 at {name} "
                  f"{peak / 1e12:.0f} TF: {outs * ex / peak * 1e3:5.2f} ms  (MMA time only; + 6 operand re-splits, 2 twiddle passes, "
                  "1 spectrum multiply per element on CUDA cores)")
    # ---- error -------------------------------------------------------------------------------------------------
    print("error / tolerance (tolerance = 1e-5 * sum|h| * max|x|), 2 blocks per case:")
    for name, x in cases.items():
        want = np.stack([np.convolve(r, h)[K - 1:N] for r in x])
        tol = 1e-5 * np.abs(h).sum() * np.abs(x).max()
        row = []
        for terms in (3, 4, 6):
            for axis_b, sc in ((None, "tile scale"), (0, "column scale")):
                ya, yb = overlap_save_block(x[0], x[1], hspec, terms, axis_b)
                err = max(np.abs(ya - want[0]).max(), np.abs(yb - want[1]).max())
                row.append(f"x{terms} {sc}: {err / tol:7.2f}")
        print(f"  {name:42s} " + " | ".join(row))
    print("f32 FFT reference (numpy complex64 fft) for scale:")
    for name, x in cases.items():
        want = np.stack([np.convolve(r, h)[K - 1:N] for r in x])
        tol = 1e-5 * np.abs(h).sum() * np.abs(x).max()
        z = (x[0] + 1j * x[1]).astype(np.complex64)
        y = np.fft.ifft((np.fft.fft(z).astype(np.complex64) * hspec.astype(np.complex64)).astype(np.complex64)).astype(np.complex64)
        err = max(np.abs(y.real[K - 1:] - want[0]).max(), np.abs(y.imag[K - 1:] - want[1]).max())
        print(f"  {name:42s} {err / tol:7.3f}")


if __name__ == "__main__":
    main()
