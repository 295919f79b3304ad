"""CPU restatement (numpy, f64) of the overlap-save kernel's ALGORITHM (scir_b200/csrc/fir_os.cu), so that its structure is
pinned independently of the GPU: in-place decimation-in-frequency passes whose output order is digit-reversed, the spectrum
produced by the same forward passes (hence permuted identically), the mirrored decimation-in-time inverse, two real blocks
packed into one complex transform, block geometry N = L + K - 1 with the last L outputs valid, causal and anticausal
direction.  If this file and the kernel ever disagree, `tests/test_gpu_os.py` (GPU vs oracle) says which one is wrong."""
import numpy as np
import pytest


def radices(logn):
    """radix-16 passes, then an innermost radix of 16 (logn % 4 == 0) or 2^(logn % 4) -- fir_os.cu: OsGeom"""
    last = 16 if logn % 4 == 0 else 1 << (logn % 4)
    n16 = (logn - (4 if logn % 4 == 0 else logn % 4)) // 4
    return [16] * n16 + [last]


def forward_dif(z, logn):
    """in place: for sub-transforms of length np_, element i + j*m (m = np_/R): b = DFT_R(a); store b[k] * W_np^(i k) at i + k*m"""
    n = 1 << logn
    s = z.astype(np.complex128).copy()
    np_ = n
    for r in radices(logn):
        m = np_ // r
        f = np.exp(-2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        for base in range(0, n, np_):
            for i in range(m):
                idx = base + i + m * np.arange(r)
                b = f @ s[idx]
                s[idx] = b * np.exp(-2j * np.pi * i * np.arange(r) / np_)
        np_ = m
    return s


def inverse_dit(s, logn):
    """the mirrored passes, innermost first: a[k] *= conj(W_np^(i k)); then the conjugate R-point DFT (no 1/R)"""
    n = 1 << logn
    s = s.copy()
    rs = radices(logn)
    lens = []
    np_ = n
    for r in rs:
        lens.append(np_)
        np_ //= r
    for r, np_ in zip(reversed(rs), reversed(lens)):
        m = np_ // r
        f = np.exp(+2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        for base in range(0, n, np_):
            for i in range(m):
                idx = base + i + m * np.arange(r)
                a = s[idx] * np.exp(+2j * np.pi * i * np.arange(r) / np_)
                s[idx] = f @ a
    return s


def position_of_frequency(logn):
    """memory position of frequency k after the forward passes: k = k0 + r0 (k1 + r1 (k2 ...)) sits at
    k0 * (N / r0) + k1 * (N / (r0 r1)) + ...  (mixed-radix digit reversal)"""
    n = 1 << logn
    pos = np.zeros(n, np.int64)
    for k in range(n):
        rem, stride, p = k, n, 0
        for r in radices(logn):
            stride //= r
            p += (rem % r) * stride
            rem //= r
        pos[k] = p
    return pos


@pytest.mark.parametrize("logn", [8, 10, 12, 14])
def test_forward_is_a_digit_reversed_dft_and_the_inverse_undoes_it(logn):
    n = 1 << logn
    rng = np.random.RandomState(logn)
    z = rng.randn(n) + 1j * rng.randn(n)
    s = forward_dif(z, logn)
    pos = position_of_frequency(logn)
    np.testing.assert_allclose(s[pos], np.fft.fft(z), atol=1e-9 * n)
    np.testing.assert_allclose(inverse_dit(s, logn) / n, z, atol=1e-10)


@pytest.mark.parametrize("logn,k,direction", [(10, 63, +1), (12, 509, +1), (12, 640, -1), (14, 4097, +1), (14, 1025, -1)])
def test_overlap_save_pairs_blocks_and_keeps_the_last_L_outputs(logn, k, direction):
    """One complex transform filters TWO consecutive L-blocks: h real => IFFT(FFT(a + i b) H) = h*a + i h*b."""
    n = 1 << logn
    L = n - (k - 1)
    rng = np.random.RandomState(k)
    c = rng.randn(k)
    x = rng.randn(5 * L + 17)
    # the spectrum comes from the SAME forward passes (os_spectrum_kernel), scaled by 1/N
    H = forward_dif(np.r_[c, np.zeros(n - k)] / n, logn)
    if direction > 0:
        want = np.convolve(x, c)[: x.size]                              # out[i] = sum_d c[d] x[i-d], zero history
    else:
        want = np.correlate(np.r_[x, np.zeros(k - 1)], c, "valid")      # out[i] = sum_d c[d] x[i+d], zero future

    def v(i):                                                           # virtual sequence, zero outside the row
        i = np.asarray(i)
        ok = (i >= 0) & (i < x.size)
        return np.where(ok, x[np.clip(i, 0, x.size - 1)], 0.0)

    got = np.zeros(x.size)
    for pair in range((x.size + 2 * L - 1) // (2 * L)):
        i0 = pair * 2 * L
        nn = np.arange(n)
        a0 = i0 - (k - 1) if direction > 0 else i0 + L + (k - 1) - 1    # fir_os.cu: element n is v[a0 + sgn n] + i v[a0 + L + sgn n]
        z = v(a0 + direction * nn) + 1j * v(a0 + L + direction * nn)
        y = inverse_dit(forward_dif(z, logn) * H, logn)
        for n_out in range(k - 1, n):                                   # outputs n >= k-1 of either block
            o = a0 + direction * n_out
            if 0 <= o < x.size:
                got[o] = y[n_out].real
            if 0 <= o + L < x.size:
                got[o + L] = y[n_out].imag
    np.testing.assert_allclose(got, want, atol=1e-9 * np.abs(c).sum() * np.abs(x).max())


def test_block_geometry_and_dispatch_model():
    """N = 4096 up to K = 640, 16384 beyond; a block pair yields 2 (N - K + 1) outputs; the cost model of api.cu: launch_fir
    (fir_os.cu: fir_os_supported) sends config 3 to the FFT kernel and config 5's fused pass to the tensor kernel."""
    import math

    def logn_for(k):
        return 12 if k <= 640 else 14

    assert logn_for(509) == 12 and logn_for(641) == 14 and logn_for(4097) == 14
    assert (1 << 14) - (4097 - 1) == 12288 and (1 << 12) - (509 - 1) == 3588

    def t_os(k, rows, n):
        lg = logn_for(k)
        L = (1 << lg) - (k - 1)
        pairs = (-(-n // L) + 1) // 2
        per_sm = math.ceil(pairs * rows / 148)
        return 12e-6 + math.ceil(per_sm / (1 if lg == 14 else 3)) * (19.9e-6 if lg == 14 else 13.5e-6)

    def t_toep(k, tiles):
        pmax = (k - 1 + 127) // 128
        ks = sum(8 - (max(0, 128 * pb - (k - 1)) >> 4) for pb in range(pmax + 1))
        return max(20e-6, 16e-6 + math.ceil(tiles / 148) * max(3.6e-6, 3 * ks * 58e-9))

    assert t_os(4097, 256, 1 << 22) < t_toep(4097, 256 * 256)             # config 3: 5.9 ms vs 20 ms (measured 5.8 / 17.0)
    assert t_os(509, 8192, 1 << 18) > t_toep(509, 8192 * 17)              # config 5 fused: 9.2 ms vs 6.2 ms (measured 9.2 / 6.4)
    assert t_os(1025, 4096, 5000) > t_toep(1025, 4096)                    # rows shorter than a block pair stay on the tensor kernel
