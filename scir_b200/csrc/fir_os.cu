// fir_os.cu -- overlap-save FIR pass for long filters: block FFT convolution in shared memory, FP32.
//
// Why: the tcgen05 block-Toeplitz kernel (fir_toeplitz.cu) runs at the chip's deliverable MMA rate, but direct form
// costs 2K flop per output -- 25 344 executed tensor flop per output at K = 4097 (config 3), 12x above that config's
// HBM floor.  The only lever left is executing fewer flops (task 6).  Overlap-save does the same linear
// convolution with O(log N) work per output: a block of N = L + K - 1 samples is transformed, multiplied by the filter's
// spectrum and transformed back; the last L outputs of the block are exact linear-convolution outputs.  The costing
// (tools/costing_overlap_save.py, profiles/r02_overlap_save_costing.txt) compared a tensor-core DFT-by-GEMM formulation
// with a plain FP32 shared-memory FFT: the GEMM route needs 8192 executed tensor flop per output plus six operand
// re-splits per element and more shared memory than an SM has; the FP32 FFT needs ~100 flop and ~55 B of shared-memory
// traffic per output and no precision split at all -- so it is the one built.
//
// Semantics are those of every other FIR pass (common.cuh: FirPass, causal direction): out[i] = sum_d c[d] * v[i - d]
// over the virtual sequence v (offset, zero / held boundary, odd / even / constant extension), which is what
// crates/scir-gpu/src/lib.rs:1141-1148 computes for the plain zero-state case.
//
// Kernel: one CTA per PAIR of consecutive L-blocks of a row.  h is real, so the two blocks ride through ONE complex
// FFT as z = a + i b:  IFFT(FFT(z) H) = (h * a) + i (h * b)  -- no unpacking step.  The transform is an in-place
// decimation-in-frequency FFT (radix 16, 16, 16 [, 4]) whose output order is digit-reversed; H is produced by the SAME
// forward transform (os_spectrum_kernel), so it is digit-reversed identically and the inverse -- the mirrored
// decimation-in-time passes -- needs no reordering either.  The first forward pass gathers straight from global memory,
// the innermost forward pass, the multiplication by H and the innermost inverse pass are one register-resident step,
// and the last inverse pass scatters straight to global memory: four shared-memory round trips for N = 4096.
// Arithmetic is FP32 throughout; error ~ log2(N) * 2^-24 relative to the block's RMS (measured 0.002-0.02 of the
// path's tolerance 1e-5 * sum|h| * max|x|).  Like the tensor kernel, a block that holds a NaN / Inf sample is
// flagged and recomputed by a fix-up kernel the reference's way, so non-finite values stay local to their K outputs.
#include "common.cuh"

#include <cmath>
#include <cstring>

namespace scir_b200 {

namespace {

constexpr int kOsTwN = 16384;                              // twiddle table: W_16384^t, t in [0, 16384)

struct OsParams {
    FirPass p;
    int k;
    long long L;                  // outputs per block
    long long pairs_per_row;
    long long total_pairs;
    const float2* H;              // [N] spectrum / N, in the forward transform's (digit-reversed) order
    const float2* tw;             // [kOsTwN]
    const float* taps;            // [k] by delay index (device): spectrum kernel and fix-up
    int* flags;                   // [total_pairs]
    int fast_ok;                  // rows may be read without the virtual-sequence rules in the interior
};

__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }      // one pad slot per 16: conflict-free stride-16 access

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// cos / sin of 2 pi k / 16
__device__ constexpr float kC16[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                      0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
__device__ constexpr float kS16[8] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                                      1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};

// R-point DFT in registers, natural order in and out (recursive decimation in time).  INV: conjugate kernel (no 1/R).
template <int R, bool INV>
__device__ __forceinline__ void dft(float2 (&a)[R])
{
    if constexpr (R == 2) {
        const float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    } else {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
            e[j] = a[2 * j];
            o[j] = a[2 * j + 1];
        }
        dft<R / 2, INV>(e);
        dft<R / 2, INV>(o);
#pragma unroll
        for (int kk = 0; kk < R / 2; ++kk) {
            float2 t;
            if (kk == 0) {
                t = o[0];
            } else if (kk * 4 == R) {                       // W = -i (forward) / +i (inverse)
                t = INV ? make_float2(-o[kk].y, o[kk].x) : make_float2(o[kk].y, -o[kk].x);
            } else {
                const float c = kC16[kk * (16 / R)], s = kS16[kk * (16 / R)];
                const float2 w = make_float2(c, INV ? s : -s);
                t = cmul(o[kk], w);
            }
            a[kk] = cadd(e[kk], t);
            a[kk + R / 2] = csub(e[kk], t);
        }
    }
}

// a[k] *= w^k (forward) or conj(w)^k (inverse), k = 1 .. R-1.  Powers come from a product tree of depth <= 4
// (w^k = w^hi * w^(k - hi), hi the top bit of k), generated in the order they are consumed so that only the
// powers of two stay live.
__host__ __device__ constexpr int os_hibit(int k) { return (k >= 8) ? 8 : (k >= 4) ? 4 : (k >= 2) ? 2 : 1; }

template <int R, bool INV>
__device__ __forceinline__ void twiddle(float2 (&a)[R], float2 w1)
{
    if (INV) w1.y = -w1.y;
    float2 w[R];
    w[1] = w1;
    a[1] = cmul(a[1], w1);
#pragma unroll
    for (int kk = 2; kk < R; ++kk) {
        const int hi = os_hibit(kk);
        w[kk] = (kk == hi) ? cmul(w[kk / 2], w[kk / 2]) : cmul(w[hi], w[kk - hi]);
        a[kk] = cmul(a[kk], w[kk]);
    }
}

template <int LOGN>
struct OsGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int THREADS = (LOGN >= 14) ? N / 32 : N / 16;       // radix-16 butterflies per thread per pass: 2 / 1
    static constexpr int MIN_CTAS = (LOGN >= 14) ? 1 : 3;
    static constexpr int RLAST = (LOGN % 4 == 0) ? 16 : (1 << (LOGN % 4));      // innermost radix: 16 (N = 4096), 4 (N = 16384)
    static constexpr int NP16 = (LOGN - ((LOGN % 4 == 0) ? 4 : (LOGN % 4))) / 4; // radix-16 passes before the innermost one
    static constexpr size_t SMEM = static_cast<size_t>(N + N / 16) * sizeof(float2);
};

// virtual input sequence (the rules of fir_direct.cu: vload)
__device__ __forceinline__ float os_vload(const FirPass& p, const float* __restrict__ xr, long long i)
{
    if (i < 0) {
        if (p.bound == BOUND_ZERO) return 0.f;
        i = 0;
    } else if (i >= p.n_v) {
        if (p.bound == BOUND_ZERO) return 0.f;
        i = p.n_v - 1;
    }
    const long long u = i + p.in_off;
    if (p.ext_mode == EXT_NONE) return xr[u];
    const long long last = p.n_x - 1;
    if (u < 0) {
        if (p.ext_mode == EXT_ODD) return 2.f * xr[0] - xr[-u];
        if (p.ext_mode == EXT_EVEN) return xr[-u];
        return xr[0];
    }
    if (u > last) {
        if (p.ext_mode == EXT_ODD) return 2.f * xr[last] - xr[2 * last - u];
        if (p.ext_mode == EXT_EVEN) return xr[2 * last - u];
        return xr[last];
    }
    return xr[u];
}

// One radix-16 forward pass on shared memory: sub-transforms of length np, m = np / 16.
template <int LOGN>
__device__ __forceinline__ void fwd_pass16(float2* s, const float2* __restrict__ tw, int lognp, int tid)
{
    constexpr int N = 1 << LOGN;
    const int logm = lognp - 4, m = 1 << logm;
    for (int q = tid; q < N / 16; q += OsGeom<LOGN>::THREADS) {
        const int i = q & (m - 1), base = ((q >> logm) << lognp) + i;
        float2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = s[pidx(base + (j << logm))];
        dft<16, false>(a);
        twiddle<16, false>(a, __ldg(tw + (static_cast<long long>(i) << (14 - lognp))));     // W_np^i = W_16384^(i * 16384 / np)
#pragma unroll
        for (int j = 0; j < 16; ++j) s[pidx(base + (j << logm))] = a[j];
    }
}

template <int LOGN>
__device__ __forceinline__ void inv_pass16(float2* s, const float2* __restrict__ tw, int lognp, int tid)
{
    constexpr int N = 1 << LOGN;
    const int logm = lognp - 4, m = 1 << logm;
    for (int q = tid; q < N / 16; q += OsGeom<LOGN>::THREADS) {
        const int i = q & (m - 1), base = ((q >> logm) << lognp) + i;
        float2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = s[pidx(base + (j << logm))];
        twiddle<16, true>(a, __ldg(tw + (static_cast<long long>(i) << (14 - lognp))));
        dft<16, true>(a);
#pragma unroll
        for (int j = 0; j < 16; ++j) s[pidx(base + (j << logm))] = a[j];
    }
}

// The innermost step: forward radix-RL pass (m = 1, no twiddles), multiplication by the spectrum, inverse radix-RL pass.
template <int LOGN, bool SPECTRUM_ONLY>
__device__ __forceinline__ void middle(float2* s, const float2* __restrict__ H, float2* __restrict__ H_out, int tid)
{
    constexpr int N = 1 << LOGN, RL = OsGeom<LOGN>::RLAST;
    for (int q = tid; q < N / RL; q += OsGeom<LOGN>::THREADS) {
        const int base = q * RL;
        float2 a[RL];
#pragma unroll
        for (int j = 0; j < RL; ++j) a[j] = s[pidx(base + j)];
        dft<RL, false>(a);
        if constexpr (SPECTRUM_ONLY) {
#pragma unroll
            for (int j = 0; j < RL; ++j) H_out[base + j] = a[j];
        } else {
#pragma unroll
            for (int j = 0; j < RL; ++j) a[j] = cmul(a[j], __ldg(H + base + j));
            dft<RL, true>(a);
#pragma unroll
            for (int j = 0; j < RL; ++j) s[pidx(base + j)] = a[j];
        }
    }
}

// Spectrum of the taps in the transform's own output order, scaled by 1/N: H[pos] = FFT(c zero-padded)[rev(pos)] / N.
template <int LOGN>
__global__ void __launch_bounds__(OsGeom<LOGN>::THREADS) os_spectrum_kernel(const float* __restrict__ taps, int k, const float2* __restrict__ tw,
                                                                            float2* __restrict__ H)
{
    extern __shared__ __align__(16) float2 s_os[];
    constexpr int N = 1 << LOGN;
    const int tid = threadIdx.x;
    const float inv_n = 1.f / static_cast<float>(N);
    for (int n = tid; n < N; n += OsGeom<LOGN>::THREADS) s_os[pidx(n)] = make_float2(n < k ? taps[n] * inv_n : 0.f, 0.f);
    __syncthreads();
    int lognp = LOGN;
#pragma unroll
    for (int pass = 0; pass < OsGeom<LOGN>::NP16; ++pass) {
        fwd_pass16<LOGN>(s_os, tw, lognp, tid);
        __syncthreads();
        lognp -= 4;
    }
    middle<LOGN, true>(s_os, nullptr, H, tid);
}

template <int LOGN>
__global__ void __launch_bounds__(OsGeom<LOGN>::THREADS, OsGeom<LOGN>::MIN_CTAS) fir_os_kernel(const __grid_constant__ OsParams q)
{
    extern __shared__ __align__(16) float2 s_os[];
    constexpr int N = 1 << LOGN;
    constexpr int NT = OsGeom<LOGN>::THREADS;
    const FirPass& p = q.p;
    const int tid = threadIdx.x;
    const long long pair = blockIdx.x;
    const long long row = pair / q.pairs_per_row;
    const long long pr = pair - row * q.pairs_per_row;
    const long long i0 = p.out_begin + pr * 2 * q.L;        // first output (virtual index) of block a; block b starts L later
    // Element n of the block is v[a0 + sgn n] + i v[a0 + L + sgn n], and output n (n >= k-1) of block a is out[a0 + sgn n]:
    //   causal     (out[i] = sum c[d] v[i-d]):  a0 = i0 - (k-1),        sgn = +1
    //   anticausal (out[i] = sum c[d] v[i+d]):  a0 = i0 + L + (k-1) - 1, sgn = -1   (the block runs backwards in time)
    const int sgn = (p.dir > 0) ? 1 : -1;
    const long long a0 = (p.dir > 0) ? i0 - (q.k - 1) : i0 + q.L + (q.k - 1) - 1;
    const float* __restrict__ xr = p.x + row * p.ld_x;
    const float2* __restrict__ tw = q.tw;

    // ---- first forward pass straight from global memory ----
    // interior blocks (everything inside the row, no extension rule applies) read x directly
    const long long v_lo = (p.dir > 0) ? a0 : a0 - (N - 1), v_hi = (p.dir > 0) ? a0 + q.L + N - 1 : a0 + q.L;
    const bool interior = q.fast_ok && v_lo >= 0 && v_hi < p.n_v && v_lo + p.in_off >= 0 && v_hi + p.in_off < p.n_x;
    int bad = 0;
    {
        constexpr int logm = LOGN - 4, m = 1 << logm;
        for (int i = tid; i < m; i += NT) {
            float2 a[16];
            if (interior) {
                const float* __restrict__ xa = xr + (a0 + p.in_off) + sgn * i;
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = make_float2(__ldg(xa + sgn * (j << logm)), __ldg(xa + q.L + sgn * (j << logm)));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const long long ia = a0 + sgn * static_cast<long long>(i + (j << logm));
                    a[j] = make_float2(os_vload(p, xr, ia), os_vload(p, xr, ia + q.L));
                }
            }
            float chk = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) chk += a[j].x * 0.f + a[j].y * 0.f;        // NaN iff any sample is NaN / Inf
            bad |= (chk != chk);
            dft<16, false>(a);
            twiddle<16, false>(a, __ldg(tw + (static_cast<long long>(i) << (14 - LOGN))));
#pragma unroll
            for (int j = 0; j < 16; ++j) s_os[pidx(i + (j << logm))] = a[j];
        }
    }
    const int any_bad = __syncthreads_or(bad);              // also the barrier after pass 0
    if (tid == 0) q.flags[pair] = any_bad;

    int lognp = LOGN - 4;
#pragma unroll
    for (int pass = 1; pass < OsGeom<LOGN>::NP16; ++pass) {
        fwd_pass16<LOGN>(s_os, tw, lognp, tid);
        __syncthreads();
        lognp -= 4;
    }
    middle<LOGN, false>(s_os, q.H, nullptr, tid);
    __syncthreads();
#pragma unroll
    for (int pass = OsGeom<LOGN>::NP16 - 1; pass >= 1; --pass) {
        lognp += 4;
        inv_pass16<LOGN>(s_os, tw, lognp, tid);
        __syncthreads();
    }

    // ---- last inverse pass straight to global memory: outputs n in [k-1, N) of either block ----
    {
        constexpr int logm = LOGN - 4, m = 1 << logm;
        float* __restrict__ yr = p.y + row * p.ld_y + p.out_off;
        for (int i = tid; i < m; i += NT) {
            float2 a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = s_os[pidx(i + (j << logm))];
            twiddle<16, true>(a, __ldg(tw + (static_cast<long long>(i) << (14 - LOGN))));
            dft<16, true>(a);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = i + (j << logm);
                if (n >= q.k - 1) {
                    const long long ia = a0 + sgn * static_cast<long long>(n);
                    if (ia < p.out_end) yr[ia] = a[j].x;
                    if (ia + q.L < p.out_end) yr[ia + q.L] = a[j].y;
                }
            }
        }
    }
}

// =====================================================================================================================
// N = 16384, packed: the same transform with TWO butterflies per thread riding in the two lanes of the packed FP32
// instructions (FADD2 / FMUL2 / FFMA2: 64 lane-operations per issue slot).  The scalar kernel above is issue-bound
// (ncu, config 3: 4.2 G warp instructions, 64 % of them FP32 arithmetic, issue slots 63 % busy at one CTA per SM); here
// every complex add is 2 instructions for two butterflies instead of 4, every complex multiply 4-5 instead of 8.  For
// that the block lives in shared memory as two PLANES (re[], im[]) and a thread owns the butterflies of two adjacent
// positions i, i+1: one LDS.64 / STS.64 fetches the same element of both -- already in packed form, no lane shuffles.
// Pass structure, index arithmetic, twiddles and the spectrum's order are exactly those of fir_os_kernel<14>.
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b)
{
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct C2 {                       // two complex numbers: (re.lo + i im.lo) and (re.hi + i im.hi)
    u64 re, im;
};
__device__ __forceinline__ C2 c2add(C2 a, C2 b) { return C2{add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ C2 c2sub(C2 a, C2 b) { return C2{sub2(a.re, b.re), sub2(a.im, b.im)}; }
__device__ __forceinline__ C2 c2mul(C2 a, C2 w)               // per-lane complex product a * w
{
    return C2{sub2(mul2(a.re, w.re), mul2(a.im, w.im)), fma2(a.re, w.im, mul2(a.im, w.re))};
}
__device__ __forceinline__ C2 c2mulc(C2 a, C2 w)              // a * conj(w)
{
    return C2{fma2(a.im, w.im, mul2(a.re, w.re)), sub2(mul2(a.im, w.re), mul2(a.re, w.im))};
}

template <int R, bool INV>
__device__ __forceinline__ void dft2x(C2 (&a)[R])
{
    if constexpr (R == 2) {
        const C2 t = a[0];
        a[0] = c2add(t, a[1]);
        a[1] = c2sub(t, a[1]);
    } else {
        C2 e[R / 2], o[R / 2];
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
            e[j] = a[2 * j];
            o[j] = a[2 * j + 1];
        }
        dft2x<R / 2, INV>(e);
        dft2x<R / 2, INV>(o);
#pragma unroll
        for (int kk = 0; kk < R / 2; ++kk) {
            if (kk == 0) {
                a[0] = c2add(e[0], o[0]);
                a[R / 2] = c2sub(e[0], o[0]);
            } else if (kk * 4 == R) {
                // t = -i o (forward) = (o.im, -o.re);  +i o (inverse) = (-o.im, o.re): folded into the add / sub pattern
                if (!INV) {
                    a[kk] = C2{add2(e[kk].re, o[kk].im), sub2(e[kk].im, o[kk].re)};
                    a[kk + R / 2] = C2{sub2(e[kk].re, o[kk].im), add2(e[kk].im, o[kk].re)};
                } else {
                    a[kk] = C2{sub2(e[kk].re, o[kk].im), add2(e[kk].im, o[kk].re)};
                    a[kk + R / 2] = C2{add2(e[kk].re, o[kk].im), sub2(e[kk].im, o[kk].re)};
                }
            } else {
                const float c = kC16[kk * (16 / R)], sn = kS16[kk * (16 / R)];
                const u64 cc = pk(c, c), ps = pk(sn, sn), ns = pk(-sn, -sn);
                // forward w = c - i s: t.re = o.re c + o.im s, t.im = o.im c - o.re s;  inverse w = c + i s
                const C2 t = INV ? C2{fma2(o[kk].im, ns, mul2(o[kk].re, cc)), fma2(o[kk].re, ps, mul2(o[kk].im, cc))}
                                 : C2{fma2(o[kk].im, ps, mul2(o[kk].re, cc)), fma2(o[kk].re, ns, mul2(o[kk].im, cc))};
                a[kk] = c2add(e[kk], t);
                a[kk + R / 2] = c2sub(e[kk], t);
            }
        }
    }
}

// a[k] *= w^k (forward) or conj(w^k) (inverse); w holds the two lanes' own twiddles
template <bool INV>
__device__ __forceinline__ void twiddle2x(C2 (&a)[16], C2 w1)
{
    C2 w[16];
    w[1] = w1;
    a[1] = INV ? c2mulc(a[1], w1) : c2mul(a[1], w1);
#pragma unroll
    for (int kk = 2; kk < 16; ++kk) {
        const int hi = os_hibit(kk);
        w[kk] = (kk == hi) ? c2mul(w[kk / 2], w[kk / 2]) : c2mul(w[hi], w[kk - hi]);
        a[kk] = INV ? c2mulc(a[kk], w[kk]) : c2mul(a[kk], w[kk]);
    }
}

constexpr int kP14N = 16384;
constexpr int kP14Threads = 512;                           // one butterfly PAIR per thread per radix-16 pass
__device__ __forceinline__ int ppad(int i) { return i + ((i >> 6) << 2); }          // 4 pad floats per 64: see fir_os.cu notes
constexpr int kP14Plane = kP14N + (kP14N >> 6) * 4;       // floats per plane
constexpr size_t kP14Smem = 2 * static_cast<size_t>(kP14Plane) * sizeof(float);

__device__ __forceinline__ C2 tw_pair(const float2* __restrict__ tw, int i0, int i1)
{
    const float2 t0 = __ldg(tw + i0), t1 = __ldg(tw + i1);
    return C2{pk(t0.x, t1.x), pk(t0.y, t1.y)};
}

// the two lanes' twiddles W_np^i, W_np^(i+1) of a pass; fetched by the caller BEFORE the barrier that precedes the pass, so
// the L2 round trip (the 128 KB table does not stay in L1 next to 139 KB of shared memory) hides behind the barrier wait
template <int LOGNP>
__device__ __forceinline__ C2 pass_twiddle(const float2* __restrict__ tw, int u)
{
    constexpr int m = 1 << (LOGNP - 4);
    const int i = (2 * u) & (m - 1);
    return tw_pair(tw, i << (14 - LOGNP), (i + 1) << (14 - LOGNP));
}

template <int LOGNP, bool INV>
__device__ __forceinline__ void pass16_2x(float* sre, float* sim, const C2 w1, int u)
{
    constexpr int logm = LOGNP - 4, m = 1 << logm;
    const int q = 2 * u;                                   // butterflies q, q + 1: adjacent positions of the same sub-transform
    const int i = q & (m - 1), base = ((q >> logm) << LOGNP) + i;
    C2 a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int pi = ppad(base + (j << logm));
        a[j].re = *reinterpret_cast<const u64*>(sre + pi);
        a[j].im = *reinterpret_cast<const u64*>(sim + pi);
    }
    if (!INV) {
        dft2x<16, false>(a);
        twiddle2x<false>(a, w1);
    } else {
        twiddle2x<true>(a, w1);
        dft2x<16, true>(a);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int pi = ppad(base + (j << logm));
        *reinterpret_cast<u64*>(sre + pi) = a[j].re;
        *reinterpret_cast<u64*>(sim + pi) = a[j].im;
    }
}

__global__ void __launch_bounds__(kP14Threads, 1) fir_os14_packed_kernel(const __grid_constant__ OsParams q)
{
    extern __shared__ __align__(16) float s_pl[];
    float* const sre = s_pl;
    float* const sim = s_pl + kP14Plane;
    constexpr int N = kP14N;
    const FirPass& p = q.p;
    const int u = threadIdx.x;
    const long long pair = blockIdx.x;
    const long long row = pair / q.pairs_per_row;
    const long long pr = pair - row * q.pairs_per_row;
    const long long i0 = p.out_begin + pr * 2 * q.L;
    const int sgn = (p.dir > 0) ? 1 : -1;
    const long long a0 = (p.dir > 0) ? i0 - (q.k - 1) : i0 + q.L + (q.k - 1) - 1;
    const float* __restrict__ xr = p.x + row * p.ld_x;
    const float2* __restrict__ tw = q.tw;
    const long long v_lo = (p.dir > 0) ? a0 : a0 - (N - 1), v_hi = (p.dir > 0) ? a0 + q.L + N - 1 : a0 + q.L;
    const bool interior = q.fast_ok && v_lo >= 0 && v_hi < p.n_v && v_lo + p.in_off >= 0 && v_hi + p.in_off < p.n_x;

    // ---- pass 0 (sub-transform length N, m = 1024) straight from global memory; lanes = positions 2u, 2u + 1 ----
    int bad;
    {
        constexpr int logm = 10;
        const int i = 2 * u;
        C2 a[16];
        if (interior) {
            const float* __restrict__ xa = xr + (a0 + p.in_off) + sgn * i;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float* e = xa + sgn * (j << logm);
                a[j].re = pk(__ldg(e), __ldg(e + sgn));
                a[j].im = pk(__ldg(e + q.L), __ldg(e + q.L + sgn));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const long long ia = a0 + sgn * static_cast<long long>(i + (j << logm));
                a[j].re = pk(os_vload(p, xr, ia), os_vload(p, xr, ia + sgn));
                a[j].im = pk(os_vload(p, xr, ia + q.L), os_vload(p, xr, ia + q.L + sgn));
            }
        }
        u64 chk = 0ull;                                    // NaN in either lane iff any sample is NaN / Inf
#pragma unroll
        for (int j = 0; j < 16; ++j) chk = fma2(a[j].re, 0ull, fma2(a[j].im, 0ull, chk));
        float c0, c1;
        upk(chk, c0, c1);
        bad = (c0 != c0) | (c1 != c1);
        dft2x<16, false>(a);
        twiddle2x<false>(a, tw_pair(tw, i, i + 1));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int pi = ppad(i + (j << logm));
            *reinterpret_cast<u64*>(sre + pi) = a[j].re;
            *reinterpret_cast<u64*>(sim + pi) = a[j].im;
        }
    }
    C2 wn = pass_twiddle<10>(tw, u);
    const int any_bad = __syncthreads_or(bad);
    if (u == 0) q.flags[pair] = any_bad;

    pass16_2x<10, false>(sre, sim, wn, u);
    wn = pass_twiddle<6>(tw, u);
    __syncthreads();
    pass16_2x<6, false>(sre, sim, wn, u);
    // the spectrum of the first innermost butterfly is requested before the barrier, each next one a trip ahead
    float4 h01 = __ldg(reinterpret_cast<const float4*>(q.H + 4 * u));
    float4 h23 = __ldg(reinterpret_cast<const float4*>(q.H + 4 * u) + 1);
    __syncthreads();
    // ---- innermost step: radix-4 forward, times the spectrum, radix-4 inverse (scalar; 4 contiguous positions) ----
#pragma unroll 2
    for (int b4 = u; b4 < N / 4; b4 += kP14Threads) {
        const int pi = ppad(4 * b4);
        const float4 r4 = *reinterpret_cast<const float4*>(sre + pi), i4 = *reinterpret_cast<const float4*>(sim + pi);
        float2 a[4] = {make_float2(r4.x, i4.x), make_float2(r4.y, i4.y), make_float2(r4.z, i4.z), make_float2(r4.w, i4.w)};
        dft<4, false>(a);
        const float4 c01 = h01, c23 = h23;
        if (b4 + kP14Threads < N / 4) {
            h01 = __ldg(reinterpret_cast<const float4*>(q.H + 4 * (b4 + kP14Threads)));
            h23 = __ldg(reinterpret_cast<const float4*>(q.H + 4 * (b4 + kP14Threads)) + 1);
        }
        a[0] = cmul(a[0], make_float2(c01.x, c01.y));
        a[1] = cmul(a[1], make_float2(c01.z, c01.w));
        a[2] = cmul(a[2], make_float2(c23.x, c23.y));
        a[3] = cmul(a[3], make_float2(c23.z, c23.w));
        dft<4, true>(a);
        *reinterpret_cast<float4*>(sre + pi) = make_float4(a[0].x, a[1].x, a[2].x, a[3].x);
        *reinterpret_cast<float4*>(sim + pi) = make_float4(a[0].y, a[1].y, a[2].y, a[3].y);
    }
    wn = pass_twiddle<6>(tw, u);
    __syncthreads();
    pass16_2x<6, true>(sre, sim, wn, u);
    wn = pass_twiddle<10>(tw, u);
    __syncthreads();
    pass16_2x<10, true>(sre, sim, wn, u);
    wn = tw_pair(tw, 2 * u, 2 * u + 1);
    __syncthreads();

    // ---- last inverse pass straight to global memory ----
    {
        constexpr int logm = 10;
        const int i = 2 * u;
        float* __restrict__ yr = p.y + row * p.ld_y + p.out_off;
        C2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int pi = ppad(i + (j << logm));
            a[j].re = *reinterpret_cast<const u64*>(sre + pi);
            a[j].im = *reinterpret_cast<const u64*>(sim + pi);
        }
        twiddle2x<true>(a, wn);
        dft2x<16, true>(a);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float ra, rb, ia_, ib;
            upk(a[j].re, ra, rb);
            upk(a[j].im, ia_, ib);
#pragma unroll
            for (int lane = 0; lane < 2; ++lane) {
                const int n = i + lane + (j << logm);
                if (n >= q.k - 1) {
                    const long long o = a0 + sgn * static_cast<long long>(n);
                    if (o < p.out_end) yr[o] = lane ? rb : ra;
                    if (o + q.L < p.out_end) yr[o + q.L] = lane ? ib : ia_;
                }
            }
        }
    }
}

// Blocks that held a NaN / Inf sample, recomputed like the reference loop (lib.rs:1138-1150): one FMA chain per output.
__global__ void __launch_bounds__(256) os_fixup_kernel(const __grid_constant__ OsParams q)
{
    const FirPass& p = q.p;
    for (long long base = static_cast<long long>(blockIdx.x) * 256; base < q.total_pairs; base += static_cast<long long>(gridDim.x) * 256) {
        const long long mine = base + threadIdx.x;
        const int flag = (mine < q.total_pairs) ? q.flags[mine] : 0;
        if (!__syncthreads_or(flag)) continue;
        __shared__ int hit[256];
        hit[threadIdx.x] = flag;
        __syncthreads();
        for (int j = 0; j < 256; ++j) {
            if (!hit[j]) continue;
            const long long pair = base + j;
            const long long row = pair / q.pairs_per_row, pr = pair - row * q.pairs_per_row;
            const long long i0 = p.out_begin + pr * 2 * q.L;
            const float* __restrict__ xr = p.x + row * p.ld_x;
            float* __restrict__ yr = p.y + row * p.ld_y + p.out_off;
            for (long long o = threadIdx.x; o < 2 * q.L; o += blockDim.x) {
                const long long i = i0 + o;
                if (i >= p.out_end) break;
                float acc = 0.f;
                for (int d = 0; d < q.k; ++d) acc = fmaf(__ldg(q.taps + d), os_vload(p, xr, p.dir > 0 ? i - d : i + d), acc);
                yr[i] = acc;
            }
        }
        __syncthreads();
    }
}

int os_logn_for(int64_t k)
{
    // cost per output ~ instructions per transform / L.  Measured (profiles/README.md): 4.3 ps per output with N = 4096
    // blocks at K = 509 (L = 3588) against 5.4 ps with N = 16384 at K = 4097 (L = 12288); scaled by N / L the larger
    // block wins from K ~ 650 on
    if (k <= 640) return 12;
    return 14;
}

}  // namespace

bool fir_os_supported(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k, double* est_seconds)
{
    if (k < 2 || k > SCIR_B200_MAX_TAPS) return false;
    if (pass.batch <= 0 || pass.out_end <= pass.out_begin) return false;
    const int logn = os_logn_for(k);
    const long long L = (1LL << logn) - (k - 1);
    if (L < (1LL << logn) / 4) return false;
    if (OsGeom<14>::SMEM > static_cast<size_t>(ctx->max_smem_optin)) return false;
    const long long blocks = (pass.out_end - pass.out_begin + L - 1) / L;
    const long long pairs = (blocks + 1) / 2;
    if (pairs * pass.batch > 0x7fffffffLL) return false;
    if (est_seconds) {
        // measured per block pair and SM (profiles/README.md): N = 16384, one CTA per SM: 19.9 us (config 3: 5.81 ms for 295.8
        // pairs per SM); N = 4096, three CTAs per SM, 13.5 us each (config 5 arm: 9.24 ms for 2048 pairs per SM).  A row shorter than
        // a pair still pays for the whole pair, which is what keeps short rows off this path.
        const double t_cta = (logn == 14) ? 19.9e-6 : 13.5e-6;
        const double resident = (logn == 14) ? 1.0 : 3.0;
        const double per_sm = std::ceil(static_cast<double>(pairs * pass.batch) / static_cast<double>(ctx->sm_count));
        *est_seconds = 12e-6 + std::ceil(per_sm / resident) * t_cta;
    }
    return true;
}

int launch_fir_os(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k)
{
    if (pass.batch <= 0 || pass.out_end <= pass.out_begin) return SCIR_B200_OK;
    if (!fir_os_supported(ctx, pass, k)) return set_error(SCIR_B200_ERR_UNSUPPORTED, "overlap-save path cannot serve k=%lld", (long long)k);
    SCIR_TRY(ctx_bind(ctx));
    const int logn = os_logn_for(k);
    const long long N = 1LL << logn;

    // twiddle table W_16384^t (once per ctx, f64 on the host)
    if (!ctx->os_tw.ptr) {
        std::vector<float2> tw(kOsTwN);
        for (int t = 0; t < kOsTwN; ++t) {
            const double ang = -2.0 * 3.14159265358979323846 * static_cast<double>(t) / static_cast<double>(kOsTwN);
            tw[static_cast<size_t>(t)] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
        }
        SCIR_TRY(ctx_scratch(ctx, ctx->os_tw, tw.size() * sizeof(float2)));
        SCIR_CUDA(cudaMemcpyAsync(ctx->os_tw.ptr, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream),
                  "cudaMemcpyAsync(twiddles)");
    }
    SCIR_TRY(ctx_scratch(ctx, ctx->os_taps, static_cast<size_t>(SCIR_B200_MAX_TAPS) * sizeof(float)));
    SCIR_TRY(ctx_scratch(ctx, ctx->os_H, static_cast<size_t>(kOsTwN) * sizeof(float2)));

    using SpecKern = void (*)(const float*, int, const float2*, float2*);
    using FirKern = void (*)(const OsParams);
    const SpecKern spec = (logn == 12) ? os_spectrum_kernel<12> : os_spectrum_kernel<14>;
    const bool packed = (logn == 14) && ctx->opt.os_packed != 0;            // two butterflies per thread in FADD2 / FFMA2 lanes
    const FirKern kern = (logn == 12) ? fir_os_kernel<12> : packed ? fir_os14_packed_kernel : fir_os_kernel<14>;
    const size_t spec_smem = (logn == 12) ? OsGeom<12>::SMEM : OsGeom<14>::SMEM;
    const size_t smem = packed ? kP14Smem : spec_smem;
    const int spec_threads = (logn == 12) ? OsGeom<12>::THREADS : OsGeom<14>::THREADS;
    const int threads = packed ? kP14Threads : spec_threads;
    static thread_local bool configured[16][3] = {};
    const int d = ctx->device & 15, li = (logn == 12) ? 0 : packed ? 2 : 1;
    if (!configured[d][li]) {
        SCIR_CUDA(cudaFuncSetAttribute(spec, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(spec_smem)), "cudaFuncSetAttribute(os_spectrum_kernel)");
        SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)), "cudaFuncSetAttribute(fir_os_kernel)");
        configured[d][li] = true;
    }
    // taps and their spectrum live in ctx-owned device buffers; redone only when the filter (or the block size) changes
    if (ctx->os_taps_host.size() != static_cast<size_t>(k) || ctx->os_logn != logn ||
        std::memcmp(ctx->os_taps_host.data(), c, static_cast<size_t>(k) * sizeof(float)) != 0) {
        ctx->os_taps_host.assign(c, c + k);
        ctx->os_logn = logn;
        SCIR_CUDA(cudaMemcpyAsync(ctx->os_taps.ptr, ctx->os_taps_host.data(), static_cast<size_t>(k) * sizeof(float), cudaMemcpyHostToDevice,
                                  ctx->stream),
                  "cudaMemcpyAsync(overlap-save taps)");
        spec<<<1, spec_threads, spec_smem, ctx->stream>>>(static_cast<const float*>(ctx->os_taps.ptr), static_cast<int>(k),
                                                 static_cast<const float2*>(ctx->os_tw.ptr), static_cast<float2*>(ctx->os_H.ptr));
        SCIR_CUDA(cudaGetLastError(), "os_spectrum_kernel launch");
        ctx->launches++;
    }

    OsParams q{};
    q.p = pass;
    q.k = static_cast<int>(k);
    q.L = N - (k - 1);
    const long long blocks = (pass.out_end - pass.out_begin + q.L - 1) / q.L;
    q.pairs_per_row = (blocks + 1) / 2;
    q.total_pairs = q.pairs_per_row * pass.batch;
    q.H = static_cast<const float2*>(ctx->os_H.ptr);
    q.tw = static_cast<const float2*>(ctx->os_tw.ptr);
    q.taps = static_cast<const float*>(ctx->os_taps.ptr);
    SCIR_TRY(ctx_scratch(ctx, ctx->toep_flags, static_cast<size_t>(q.total_pairs) * sizeof(int)));
    q.flags = static_cast<int*>(ctx->toep_flags.ptr);
    q.fast_ok = 1;
    kern<<<static_cast<unsigned>(q.total_pairs), threads, smem, ctx->stream>>>(q);
    SCIR_CUDA(cudaGetLastError(), "fir_os_kernel launch");
    const long long fix_grid = std::min<long long>((q.total_pairs + 255) / 256, static_cast<long long>(ctx->sm_count) * 8);
    os_fixup_kernel<<<static_cast<unsigned>(fix_grid), 256, 0, ctx->stream>>>(q);
    SCIR_CUDA(cudaGetLastError(), "os_fixup_kernel launch");
    ctx->launches += 2;
    ctx->fixup_launches++;
    ctx->os_launches++;
    return SCIR_B200_OK;
}

}  // namespace scir_b200
