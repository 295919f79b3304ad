// upfirdn_poly.cu -- tiled polyphase up-FIR-down kernel for sm_100a (SciPy semantics, mode='constant').
//
// Spec: scipy/signal/_upfirdn_apply.pyx:421-481 / _upfirdn.py:47-64.  Output m of a row is
//     y[m] = sum_i h[t + i*up] * x[q - i],   t = (m*down) % up,  q = (m*down) / up
// so only the `1/up` of the taps that meet a real sample are ever multiplied -- the reference's
// legacy resampler (crates/scir-signal/src/lib.rs:348-360) multiplies every stuffed zero and then
// throws away 2 of 3 results.
//
// Same execution model as fir_direct.cu (FFMA-issue bound, so everything else is kept off the compute
// warps), extended to a rational rate:
//   * the phase pattern repeats every `up` outputs / `down` inputs, so a thread owns R = up*G
//     consecutive outputs starting at a multiple of `up`: phase t_j and input offset dq_j of its
//     j-th output are COMPILE-TIME constants, and the x-major sliding window is fully unrolled with
//     taps as constant-bank / uniform-register FFMA operands.
//   * per-thread input stride is down*G floats; G is chosen per (up, down) so that stride is a
//     multiple of 4 with an odd quotient: 16-B aligned, bank-conflict-free LDS.128 on a DENSE tile,
//     which is what allows the tile + halo to arrive by ONE 1-D TMA bulk copy.
//   * Z = number of leading zero taps (< up): resample_poly pads the filter in front
//     (_signaltools.py:3912-3920), which shifts which phases start at i = 1.  Phases t < Z simply
//     use i in [1, KP] instead of [0, KP): no multiplies by the structural zeros
//     (config 4: 96 taps -> exactly 32 FFMA per output).
//   * outputs go back through shared memory and one bulk store.
#include "common.cuh"

#include <memory>
#include "ptx.cuh"
#include "ext_modes.cuh"

#include <algorithm>

namespace scir_b200 {

namespace {

constexpr int kPolyNT = 256;
constexpr int kPolyMaxH = 1024;          // taps ride in the kernel parameters

struct PolyTaps {
    float c[kPolyMaxH];
};

struct PolyParams {
    const float* x;
    float* y;
    long long ld_x, ld_y;
    long long n_in;
    long long m_begin, m_end;     // outputs [m_begin, m_end) land at y[row, m - m_begin]
    long long base_m;             // first output of tile 0 (a multiple of 4*up: q0 is 16-B aligned)
    long long ntiles;
    int nchunk;
    int in_vec_ok, out_vec_ok;
    ExtSpec ext;                  // value of samples outside [0, n_in) (edge tiles only)
};

// ---- packed FP32 (FFMA2) form ------------------------------------------------------------------------
// sm_100 issues `FFMA2 Racc.F32x2, Rx.F32, URtaps.F32x2, Racc.F32x2`: two FMAs per issue slot, the sample
// broadcast, the two taps a uniform-register pair.  Measured (scir_b200_microbench_ffma2): 74.0 TFLOP/s = 99 % of
// the nominal FP32 peak against 67.7 for the scalar FFMA stream, and -- what matters at the ridge -- only HALF
// the issue slots are spent on arithmetic, so loads, stores and loop control ride along for free.
// A thread's outputs are paired (j, j+1); for every window sample s the pair's two taps (either may be a
// structural zero) are precomputed on the host into a table laid out in exactly the order the fully
// unrolled loop consumes it, so each FFMA2 takes its taps with one LDCU.64 at a compile-time offset.
constexpr int kPolyMaxPairs = 768;

struct PolyPairs {
    float2 p[kPolyMaxPairs];
};

template <int UP, int DOWN, int G, int KCP, int Z>
struct PolyGeom {
    static constexpr int R = UP * G;
    static constexpr int ZP = (Z > 0) ? 4 : 0;
    static constexpr int DQMAX = ((R - 1) * DOWN) / UP;
    static constexpr int W4 = (KCP + ZP + DQMAX + 1 + 3) / 4;
    // index into the (single-chunk) tap array of the tap that output j of a thread applies to window sample s, or -1
    __host__ __device__ static constexpr int tap_of(int j, int s)
    {
        const int tj = (j * DOWN) % UP, dq = (j * DOWN) / UP, ot = (tj < Z) ? 1 : 0;
        const int il = dq + KCP + ZP - s - ot;
        return (il >= 0 && il < KCP) ? (il + ot) * UP + tj : -1;
    }
    __host__ __device__ static constexpr int pair_count()
    {
        int n = 0;
        for (int s = 0; s < 4 * W4; ++s)
            for (int p = 0; p < R / 2; ++p)
                if (tap_of(2 * p, s) >= 0 || tap_of(2 * p + 1, s) >= 0) ++n;
        return n;
    }
    // ---- tap-reuse form (poly_core3) ---------------------------------------------------------------------------
    // Output pairs p and p + PP apply the SAME tap pair, SH samples further on: the phase pattern of a pair repeats
    // every 2*UP outputs (UP odd) or every UP outputs (UP even).  So one uniform-register tap pair feeds up to GP
    // FFMA2s instead of one, and the table holds only the pairs of the first period.
    static constexpr int NPAIR = R / 2;
    static constexpr int PP = (UP % 2 == 1) ? UP : UP / 2;
    static constexpr int SH = (UP % 2 == 1) ? 2 * DOWN : DOWN;
    static constexpr int GP = (NPAIR + PP - 1) / PP;
    static constexpr int LA = SH * (GP - 1);                   // furthest sample a period-0 pair's tap reaches ahead
    __host__ __device__ static constexpr bool pair_live(int p0, int s)
    {
        return p0 < NPAIR && (tap_of(2 * p0, s) >= 0 || tap_of(2 * p0 + 1, s) >= 0);
    }
    __host__ __device__ static constexpr int pair_count3()
    {
        int n = 0;
        for (int s = 0; s < 4 * W4; ++s)
            for (int p0 = 0; p0 < PP; ++p0)
                if (pair_live(p0, s)) ++n;
        return n;
    }
    // shift invariance the reuse rests on, and the window bound of the furthest reuse (checked at compile time)
    __host__ __device__ static constexpr bool reuse_ok()
    {
        for (int s = 0; s < 4 * W4; ++s)
            for (int p0 = 0; p0 < PP; ++p0)
                for (int g = 1; g < GP; ++g) {
                    const int p = p0 + g * PP;
                    if (p >= NPAIR) continue;
                    const int s2 = s + g * SH;
                    const int a = tap_of(2 * p0, s), b = tap_of(2 * p0 + 1, s);
                    if (s2 >= 4 * W4) {
                        if (a >= 0 || b >= 0) return false;
                        continue;
                    }
                    if (tap_of(2 * p, s2) != a || tap_of(2 * p + 1, s2) != b) return false;
                }
        // and no live (pair, sample) of a later period is missed: its period-0 image must exist inside the window
        for (int s2 = 0; s2 < 4 * W4; ++s2)
            for (int p = PP; p < NPAIR; ++p) {
                const int g = p / PP, s = s2 - g * SH;
                if ((tap_of(2 * p, s2) >= 0 || tap_of(2 * p + 1, s2) >= 0) && s < 0) return false;
            }
        return true;
    }
    static void fill_pairs3(const float* c, PolyPairs* out)    // host: same enumeration order as poly_core3
    {
        int n = 0;
        for (int s = 0; s < 4 * W4; ++s)
            for (int p0 = 0; p0 < PP; ++p0)
                if (pair_live(p0, s)) {
                    const int a = tap_of(2 * p0, s), b = tap_of(2 * p0 + 1, s);
                    out->p[n++] = make_float2(a >= 0 ? c[a] : 0.f, b >= 0 ? c[b] : 0.f);
                }
    }
    static void fill_pairs(const float* c, PolyPairs* out)     // host: same enumeration order as poly_core2
    {
        int n = 0;
        for (int s = 0; s < 4 * W4; ++s)
            for (int p = 0; p < R / 2; ++p) {
                const int a = tap_of(2 * p, s), b = tap_of(2 * p + 1, s);
                if (a >= 0 || b >= 0) out->p[n++] = make_float2(a >= 0 ? c[a] : 0.f, b >= 0 ? c[b] : 0.f);
            }
    }
};

__device__ __forceinline__ void fma2_bcast(unsigned long long& acc, float x, const float2& taps)
{
    unsigned long long xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(*reinterpret_cast<const unsigned long long*>(&taps)));
}

template <int UP, int DOWN, int G, int KCP, int Z>
__device__ __forceinline__ void poly_core2(float (&acc)[UP * G], const float* wbase, const PolyTaps& taps, const PolyPairs& pairs)
{
    using Geo = PolyGeom<UP, DOWN, G, KCP, Z>;
    constexpr int R = Geo::R;
    static_assert(Geo::pair_count() <= kPolyMaxPairs, "pair table too small");
    unsigned long long acc2[R / 2];
    float tail = 0.f;                                              // odd R: the last output stays scalar
#pragma unroll
    for (int p = 0; p < R / 2; ++p) acc2[p] = 0ull;
    const float4* w = reinterpret_cast<const float4*>(wbase - KCP - Geo::ZP);
    int cnt = 0;                                                   // compile-time after unrolling
#pragma unroll
    for (int v4 = 0; v4 < Geo::W4; ++v4) {
        const float4 v = w[v4];
        const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int s = 4 * v4 + e;
#pragma unroll
            for (int p = 0; p < R / 2; ++p) {
                if (Geo::tap_of(2 * p, s) >= 0 || Geo::tap_of(2 * p + 1, s) >= 0) {
                    fma2_bcast(acc2[p], xv[e], pairs.p[cnt]);
                    ++cnt;
                }
            }
            if constexpr (R % 2 == 1) {
                const int a = Geo::tap_of(R - 1, s);
                if (a >= 0) tail = fmaf(taps.c[a], xv[e], tail);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < R / 2; ++p)
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * p]), "=f"(acc[2 * p + 1]) : "l"(acc2[p]));
    if constexpr (R % 2 == 1) acc[R - 1] = tail;
}

// Same arithmetic as poly_core2 -- the same (output pair, sample) FFMA2s with the same taps -- but ordered so that a
// tap pair is loaded ONCE (one LDCU.64) and applied to every output pair that shares it: pairs p0, p0+PP, p0+2PP, ...
// at samples s, s+SH, s+2SH, ...  poly_core2 spends one LDCU.64 per FFMA2 (ncu, config 4: LDCU = 28 % of all issued
// instructions, one for every 1.9 FFMA2); here it is one per GP FFMA2 (config 4: GP = 5).  The sample window is
// pulled through registers by LDS.128 just ahead of its first use (LA samples of look-ahead), so the live window is
// ~LA + 8 registers instead of the whole tile row.
template <int UP, int DOWN, int G, int KCP, int Z>
__device__ __forceinline__ void poly_core3(float (&acc)[UP * G], const float* wbase, const PolyTaps& taps, const PolyPairs& pairs)
{
    using Geo = PolyGeom<UP, DOWN, G, KCP, Z>;
    constexpr int R = Geo::R, W4 = Geo::W4, PP = Geo::PP, SH = Geo::SH, GP = Geo::GP, NPAIR = Geo::NPAIR;
    static_assert(Geo::pair_count3() <= kPolyMaxPairs, "pair table too small");
    static_assert(Geo::reuse_ok(), "tap pairs are not shift-invariant across periods for this geometry");
    constexpr int LA4 = (Geo::LA + 3) / 4 + 1;                     // float4s kept loaded ahead of the current one
    unsigned long long acc2[NPAIR > 0 ? NPAIR : 1];
    float tail = 0.f;                                              // odd R: the last output stays scalar
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) acc2[p] = 0ull;
    const float4* w = reinterpret_cast<const float4*>(wbase - KCP - Geo::ZP);
    float xw[4 * W4];                                              // static indices only: lives in registers
#pragma unroll
    for (int v4 = 0; v4 < W4; ++v4) {
        if (v4 < LA4) {
            const float4 v = w[v4];
            xw[4 * v4] = v.x; xw[4 * v4 + 1] = v.y; xw[4 * v4 + 2] = v.z; xw[4 * v4 + 3] = v.w;
        }
    }
    int cnt = 0;                                                   // compile-time after unrolling
#pragma unroll
    for (int b = 0; b < W4; ++b) {
        if (b + LA4 < W4) {
            const float4 v = w[b + LA4];
            xw[4 * (b + LA4)] = v.x; xw[4 * (b + LA4) + 1] = v.y; xw[4 * (b + LA4) + 2] = v.z; xw[4 * (b + LA4) + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int s = 4 * b + e;
#pragma unroll
            for (int p0 = 0; p0 < PP; ++p0) {
                if (Geo::pair_live(p0, s)) {
                    const float2 tp = pairs.p[cnt];
                    ++cnt;
#pragma unroll
                    for (int g = 0; g < GP; ++g) {
                        if (p0 + g * PP < NPAIR && s + g * SH < 4 * W4) fma2_bcast(acc2[p0 + g * PP], xw[s + g * SH], tp);
                    }
                }
            }
            if constexpr (R % 2 == 1) {
                const int a = Geo::tap_of(R - 1, s);
                if (a >= 0) tail = fmaf(taps.c[a], xw[s], tail);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NPAIR; ++p)
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * p]), "=f"(acc[2 * p + 1]) : "l"(acc2[p]));
    if constexpr (R % 2 == 1) acc[R - 1] = tail;
}

// The FFMA core shared by the tile and the streaming kernel: R = UP*G outputs of one thread, x-major over the
// window, taps as constant-bank operands.  wbase points at the thread's sample q0_thread.
template <int UP, int DOWN, int G, int KCP, int Z, int NCH>
__device__ __forceinline__ void poly_core(float (&acc)[UP * G], const float* wbase, int nchunk, const PolyTaps& taps)
{
    constexpr int R = UP * G;
    constexpr int ZP = (Z > 0) ? 4 : 0;
    constexpr int DQMAX = ((R - 1) * DOWN) / UP;
    constexpr int W4 = (KCP + ZP + DQMAX + 1 + 3) / 4;
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.f;
    auto chunk = [&](int c) {
        // window: samples q0_thread - (c+1)*KCP - ZP + s, s in [0, 4*W4)
        const float4* w = reinterpret_cast<const float4*>(wbase - (c + 1) * KCP - ZP);
        const float* tc = taps.c + c * (KCP * UP);
#pragma unroll
        for (int v4 = 0; v4 < W4; ++v4) {
            const float4 v = w[v4];
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = 4 * v4 + e;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int tj = (j * DOWN) % UP;
                    const int dq = (j * DOWN) / UP;
                    const int ot = (tj < Z) ? 1 : 0;
                    const int il = dq + KCP + ZP - s - ot;          // chunk-local tap index of this sample
                    if (il >= 0 && il < KCP) acc[j] = fmaf(tc[(il + ot) * UP + tj], xv[e], acc[j]);
                }
            }
        }
    };
    if constexpr (NCH > 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) chunk(c);
    } else {
        for (int c = 0; c < nchunk; ++c) chunk(c);
    }
}

template <int R>
__device__ __forceinline__ void poly_store_acc(float* so, const float (&acc)[R])
{
    if constexpr (R % 4 == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 4)
            *reinterpret_cast<float4*>(so + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    } else if constexpr (R % 2 == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 2) *reinterpret_cast<float2*>(so + j) = make_float2(acc[j], acc[j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < R; ++j) so[j] = acc[j];
    }
}

// UP/DOWN: rate; G: groups of `up` outputs per thread; KCP: taps per phase per chunk; Z: leading
// zero taps; NCH: compile-time chunk count (0 = runtime loop, taps through LDCU).
template <int UP, int DOWN, int G, int KCP, int Z, int NCH>
__global__ void __launch_bounds__(kPolyNT, 3)
upfirdn_tile_kernel(const __grid_constant__ PolyParams q, const __grid_constant__ PolyTaps taps,
                    const __grid_constant__ PolyPairs pairs, int packed)
{
    constexpr int NT = kPolyNT;
    constexpr int R = UP * G;
    constexpr int SIN = DOWN * G;
    static_assert(SIN % 4 == 0 && (SIN / 4) % 2 == 1, "dense tile must be conflict-free for LDS.128");
    static_assert(KCP % 4 == 0 && Z < UP, "chunk geometry");
    constexpr int TILE_OUT = NT * R;
    constexpr int TILE_IN = NT * SIN;
    constexpr int ZP = (Z > 0) ? 4 : 0;
    constexpr int DQMAX = ((R - 1) * DOWN) / UP;
    constexpr int W4 = (KCP + ZP + DQMAX + 1 + 3) / 4;

    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long mbar;

    const int tid = threadIdx.x;
    const int nchunk = (NCH > 0) ? NCH : q.nchunk;
    const int HALO = nchunk * KCP + ZP;
    const int len = HALO + TILE_IN + 4;
    const long long bid = blockIdx.x;
    const long long row = bid / q.ntiles;
    const long long tile = bid - row * q.ntiles;
    const long long m0 = q.base_m + tile * TILE_OUT;
    const long long q0 = (m0 / UP) * DOWN;
    const long long a = q0 - HALO;                        // first input sample staged
    const float* __restrict__ xr = q.x + row * q.ld_x;
    float* __restrict__ yr = q.y + row * q.ld_y;

    const bool bulk_in = q.in_vec_ok && a >= 0 && a + len <= q.n_in;
    if (bulk_in) {
        const uint32_t bar = smem_u32(&mbar);
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(bar, static_cast<uint32_t>(len) * 4u);
            bulk_copy_g2s(smem_u32(smem), xr + a, static_cast<uint32_t>(len) * 4u, bar);
        }
        __syncthreads();
        mbar_wait(bar, 0);
    } else {
        for (int s = tid; s < len; s += NT) smem[s] = upfirdn_sample(xr, a + s, q.n_in, q.ext);
        __syncthreads();
    }

    float acc[R];
    if constexpr (NCH == 1) {
        if (packed) poly_core2<UP, DOWN, G, KCP, Z>(acc, smem + HALO + tid * SIN, taps, pairs);
        else poly_core<UP, DOWN, G, KCP, Z, NCH>(acc, smem + HALO + tid * SIN, nchunk, taps);
    } else {
        poly_core<UP, DOWN, G, KCP, Z, NCH>(acc, smem + HALO + tid * SIN, nchunk, taps);
    }

    __syncthreads();                                        // every warp is done reading the tile
    poly_store_acc<R>(smem + tid * R, acc);

    const bool bulk_out = q.out_vec_ok && m0 >= q.m_begin && m0 + TILE_OUT <= q.m_end;
    if (bulk_out) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_copy_s2g(yr + (m0 - q.m_begin), smem_u32(smem), TILE_OUT * 4u);
            bulk_store_wait_read<0>();
        }
    } else {
        __syncthreads();
        for (int s = tid; s < TILE_OUT; s += NT) {
            const long long m = m0 + s;
            if (m >= q.m_begin && m < q.m_end) yr[m - q.m_begin] = smem[s];
        }
    }
}

// ---- the streaming kernel (default): persistent CTAs, double-buffered TMA bulk loads --------------------
// Same tiles as upfirdn_tile_kernel, but a CTA walks tiles w = blockIdx.x, +gridDim.x, ...: two input
// stages are filled by cp.async.bulk two tiles ahead (one elected thread, `full` mbarriers), results leave
// through a third buffer whose bulk store drains while the next tile is being computed.  Index set-up,
// barrier init and the first DRAM round trip are paid once per CTA instead of once per tile, and a CTA
// never sits idle on its own load: at the ridge (config 4 is HBM- and FP32-bound at once) that overlap is
// the whole game.
template <int UP, int DOWN, int G, int KCP, int Z, int NCH>
__global__ void __launch_bounds__(kPolyNT, 3)
upfirdn_stream_kernel(const __grid_constant__ PolyParams q, const __grid_constant__ PolyTaps taps, long long batch,
                      const __grid_constant__ PolyPairs pairs, int packed)
{
    constexpr int NT = kPolyNT;
    constexpr int R = UP * G;
    constexpr int SIN = DOWN * G;
    static_assert(SIN % 4 == 0 && (SIN / 4) % 2 == 1, "dense tile must be conflict-free for LDS.128");
    static_assert(KCP % 4 == 0 && Z < UP, "chunk geometry");
    constexpr int TILE_OUT = NT * R;
    constexpr int TILE_IN = NT * SIN;
    constexpr int ZP = (Z > 0) ? 4 : 0;

    extern __shared__ __align__(128) float smem[];     // in[0] | in[1] | out
    __shared__ __align__(8) unsigned long long full[2];

    const int tid = threadIdx.x;
    const int nchunk = (NCH > 0) ? NCH : q.nchunk;
    const int HALO = nchunk * KCP + ZP;
    const int len = HALO + TILE_IN + 4;
    float* const out_s = smem + 2 * len;
    const uint32_t bar0 = smem_u32(&full[0]);
    const long long GD = gridDim.x;
    const long long step_row = GD / q.ntiles, step_tile = GD - step_row * q.ntiles;

    long long row = blockIdx.x / q.ntiles, tile = blockIdx.x - row * q.ntiles;
    long long nrow = row, ntile = tile;                    // runs two tiles ahead
    auto advance = [&](long long& r, long long& t) {
        r += step_row;
        t += step_tile;
        if (t >= q.ntiles) {
            t -= q.ntiles;
            r += 1;
        }
    };
    auto first_sample = [&](long long t) {
        const long long m0 = q.base_m + t * TILE_OUT;
        return (m0 / UP) * DOWN - HALO;
    };
    auto can_bulk = [&](long long a) { return q.in_vec_ok && a >= 0 && a + len <= q.n_in; };
    auto prefetch = [&](long long r, long long t, int s) { // thread 0 only
        if (r >= batch) return;
        const long long a = first_sample(t);
        if (!can_bulk(a)) return;
        mbar_arrive_expect_tx(bar0 + 8u * s, static_cast<uint32_t>(len) * 4u);
        bulk_copy_g2s(smem_u32(smem + s * len), q.x + r * q.ld_x + a, static_cast<uint32_t>(len) * 4u, bar0 + 8u * s);
    };

    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_mbar_init();
        prefetch(nrow, ntile, 0);
    }
    advance(nrow, ntile);
    if (tid == 0) prefetch(nrow, ntile, 1);
    advance(nrow, ntile);
    __syncthreads();                                       // barrier init visible to the waiters

    uint32_t phases = 0;                                   // bit s = parity to wait for on full[s]
    for (int it = 0; row < batch; ++it) {
        const int s = it & 1;
        const long long m0 = q.base_m + tile * TILE_OUT;
        const long long a = first_sample(tile);
        const float* __restrict__ xr = q.x + row * q.ld_x;
        float* __restrict__ yr = q.y + row * q.ld_y;
        float* const in = smem + s * len;

        if (can_bulk(a)) {
            mbar_wait(bar0 + 8u * s, (phases >> s) & 1u);
            phases ^= 1u << s;
        } else {                                           // edge tile: samples outside the row by extension mode
#pragma unroll 4
            for (int t = tid; t < len; t += NT) in[t] = upfirdn_sample(xr, a + t, q.n_in, q.ext);
            __syncthreads();
        }

        float acc[R];
        if constexpr (NCH == 1) {
            if (packed == 2) poly_core3<UP, DOWN, G, KCP, Z>(acc, in + HALO + tid * SIN, taps, pairs);
            else if (packed) poly_core2<UP, DOWN, G, KCP, Z>(acc, in + HALO + tid * SIN, taps, pairs);
            else poly_core<UP, DOWN, G, KCP, Z, NCH>(acc, in + HALO + tid * SIN, nchunk, taps);
        } else {
            poly_core<UP, DOWN, G, KCP, Z, NCH>(acc, in + HALO + tid * SIN, nchunk, taps);
        }

        if (tid == 0) bulk_store_wait_read<0>();           // the previous tile's store has drained `out`
        __syncthreads();                                   // A: in[s] consumed by every warp, out free
        if (tid == 0) prefetch(nrow, ntile, s);            // refill in[s] two tiles ahead

        poly_store_acc<R>(out_s + tid * R, acc);

        const bool bulk_out = q.out_vec_ok && m0 >= q.m_begin && m0 + TILE_OUT <= q.m_end;
        if (bulk_out) {
            fence_proxy_async_smem();                      // generic-proxy writes -> async proxy
            __syncthreads();                               // B
            if (tid == 0) bulk_copy_s2g(yr + (m0 - q.m_begin), smem_u32(out_s), TILE_OUT * 4u);
        } else {
            __syncthreads();
            for (int t = tid; t < TILE_OUT; t += NT) {
                const long long m = m0 + t;
                if (m >= q.m_begin && m < q.m_end) yr[m - q.m_begin] = out_s[t];
            }
        }
        advance(row, tile);
        advance(nrow, ntile);
    }
    if (tid == 0) bulk_store_wait_read<0>();               // smem must outlive the last store's read
}

// ---- the warp-specialised kernel (default for single-chunk filters) -------------------------------------------
// Same tiles and the same arithmetic as upfirdn_stream_kernel, but NO CTA barrier in the steady state:
//   * warp 8 is the producer: one lane keeps kWsStages input stages filled by cp.async.bulk (`full` mbarriers carry
//     the byte count, `empty` mbarriers are arrived on by the 8 compute warps when they are done with a stage);
//     edge tiles, whose samples come from the extension mode, are synthesised by the producer's 32 lanes;
//   * warps 0-7 compute (poly_core3: one tap-pair load per GP FFMA2s) and each drains ITS 32*R outputs with its own
//     bulk store from its own slice of the output buffer -- so a fast warp never waits for a slow one, and the only
//     things a compute warp ever waits for are its input stage and its own previous store.
// What the streaming kernel lost (ncu, config 4: 0.66 barrier stalls per issue, 29 % of the warp slots idle, loads
// and stores parked behind two __syncthreads per tile) is what this removes.
constexpr int kWsThreads = kPolyNT + 32;

template <int UP, int DOWN, int G, int KCP, int Z, int kWsStages, int CORE>
__global__ void __launch_bounds__(kWsThreads, (kWsStages >= 3) ? 2 : 3)
upfirdn_ws_kernel(const __grid_constant__ PolyParams q, const __grid_constant__ PolyTaps taps, long long batch,
                  const __grid_constant__ PolyPairs pairs)
{
    constexpr int NT = kPolyNT;
    constexpr int R = UP * G;
    constexpr int SIN = DOWN * G;
    static_assert(SIN % 4 == 0 && (SIN / 4) % 2 == 1, "dense tile must be conflict-free for LDS.128");
    constexpr int TILE_OUT = NT * R;
    constexpr int TILE_IN = NT * SIN;
    constexpr int ZP = (Z > 0) ? 4 : 0;
    constexpr int HALO = KCP + ZP;
    constexpr int LEN = HALO + TILE_IN + 4;
    constexpr int WOUT = 32 * R;                          // outputs of one compute warp
    static_assert(LEN % 4 == 0 && (WOUT * 4) % 16 == 0, "16-byte bulk copies");

    extern __shared__ __align__(128) float smem[];        // in[0..kWsStages) | out
    __shared__ __align__(8) unsigned long long bars[2 * kWsStages];     // full[s], empty[s]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* const out_s = smem + kWsStages * LEN;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kWsStages]);
    const long long GD = gridDim.x;
    const long long step_row = GD / q.ntiles, step_tile = GD - step_row * q.ntiles;
    long long row = blockIdx.x / q.ntiles, tile = blockIdx.x - row * q.ntiles;
    auto advance = [&](long long& r, long long& t) {
        r += step_row;
        t += step_tile;
        if (t >= q.ntiles) {
            t -= q.ntiles;
            r += 1;
        }
    };
    auto first_sample = [&](long long t) {
        const long long m0 = q.base_m + t * TILE_OUT;
        return (m0 / UP) * DOWN - HALO;
    };

    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, NT / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();                                       // the only CTA barrier: mbarrier init is visible

    if (warp == NT / 32) {
        // ---------------- producer ----------------
        for (int it = 0; row < batch; ++it) {
            const int s = it % kWsStages;
            const uint32_t k = static_cast<uint32_t>(it / kWsStages);
            if (k > 0) mbar_wait(empty0 + 8u * s, (k - 1) & 1u);       // the stage's previous tile has been consumed
            const long long a = first_sample(tile);
            const float* __restrict__ xr = q.x + row * q.ld_x;
            float* const in = smem + s * LEN;
            if (lane == 0) {
                if (q.in_vec_ok && a >= 0 && a + LEN <= q.n_in) {
                    mbar_arrive_expect_tx(full0 + 8u * s, LEN * 4u);
                    bulk_copy_g2s(smem_u32(in), xr + a, LEN * 4u, full0 + 8u * s);
                } else {
                    // edge tile (2 per row): its samples come from the extension mode.  The compute warps synthesise it
                    // themselves, 256 threads wide; the producer only hands the (free) stage over.  (Filling it here,
                    // 32 lanes wide, stalls the whole pipeline: measured 160 us per edge tile.)
                    mbar_arrive(full0 + 8u * s);
                }
            }
            advance(row, tile);
        }
        return;
    }

    // ---------------- compute warps ----------------
    float* const my_out = out_s + warp * WOUT;
    for (int it = 0; row < batch; ++it) {
        const int s = it % kWsStages;
        const uint32_t k = static_cast<uint32_t>(it / kWsStages);
        const long long m0 = q.base_m + tile * TILE_OUT + static_cast<long long>(warp) * WOUT;     // this warp's first output
        float* __restrict__ yr = q.y + row * q.ld_y;
        const float* const in = smem + s * LEN;

        mbar_wait(full0 + 8u * s, k & 1u);
        {
            const long long a = first_sample(tile);
            if (!(q.in_vec_ok && a >= 0 && a + LEN <= q.n_in)) {       // edge tile: see the producer
                const float* __restrict__ xr = q.x + row * q.ld_x;
                float* const inw = smem + s * LEN;
#pragma unroll 4
                for (int t = tid; t < LEN; t += NT) inw[t] = upfirdn_sample(xr, a + t, q.n_in, q.ext);
                asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // compute warps only; the producer is not part of it
            }
        }
        float acc[R];
        if constexpr (CORE == 3) poly_core3<UP, DOWN, G, KCP, Z>(acc, in + HALO + tid * SIN, taps, pairs);
        else poly_core2<UP, DOWN, G, KCP, Z>(acc, in + HALO + tid * SIN, taps, pairs);
        __syncwarp();                                      // every lane's reads of the stage are done
        if (lane == 0) {
            mbar_arrive(empty0 + 8u * s);
            bulk_store_wait_read<0>();                     // this warp's previous store has drained its slice
        }
        __syncwarp();
        poly_store_acc<R>(my_out + lane * R, acc);
        if (q.out_vec_ok && m0 >= q.m_begin && m0 + WOUT <= q.m_end) {
            fence_proxy_async_smem();                      // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) bulk_copy_s2g(yr + (m0 - q.m_begin), smem_u32(my_out), WOUT * 4u);
        } else {
            __syncwarp();
            for (int t = lane; t < WOUT; t += 32) {
                const long long m = m0 + t;
                if (m >= q.m_begin && m < q.m_end) yr[m - q.m_begin] = my_out[t];
            }
            __syncwarp();                                  // the slice is rewritten by the next tile
        }
        advance(row, tile);
    }
    if (lane == 0) bulk_store_wait_read<0>();              // smem must outlive the last store's read
}

// ---- the warp-pipelined kernel (A/B: upfirdn_variant = 6) -------------------------------------------------
// No CTA barrier anywhere: every WARP owns two input stages and one output buffer, walks warp-tiles (32*R outputs)
// with stride "all warps of the grid", prefetches two tiles ahead with cp.async.bulk on its own mbarriers and
// drains results with a bulk store.  With the packed core arithmetic takes half the issue slots, so what the
// streaming kernel still loses is its two CTA barriers per tile; here warps drift freely.  Single-chunk filters only.
constexpr int kPolyWarps = 8;

template <int UP, int DOWN, int G, int KCP, int Z>
__global__ void __launch_bounds__(kPolyWarps * 32, 3)
upfirdn_warp_kernel(const __grid_constant__ PolyParams q, const __grid_constant__ PolyTaps taps, long long batch,
                    const __grid_constant__ PolyPairs pairs, long long wtiles)
{
    constexpr int R = UP * G;
    constexpr int SIN = DOWN * G;
    constexpr int ZP = (Z > 0) ? 4 : 0;
    constexpr int HALO = KCP + ZP;
    constexpr int WT_OUT = 32 * R;
    constexpr int WT_IN = 32 * SIN;
    constexpr int LEN = HALO + WT_IN + 4;
    static_assert(LEN % 4 == 0 && WT_OUT % 4 == 0 && (WT_OUT % (4 * UP)) == 0, "16-byte bulk copies and aligned q0");

    extern __shared__ __align__(128) float smem[];     // per warp: in[0] | in[1] | out
    __shared__ __align__(8) unsigned long long full[kPolyWarps][2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* const mine = smem + warp * (2 * LEN + WT_OUT);
    float* const out_s = mine + 2 * LEN;
    const uint32_t bar0 = smem_u32(&full[warp][0]);
    const long long GW = static_cast<long long>(gridDim.x) * kPolyWarps;
    const long long step_row = GW / wtiles, step_tile = GW - step_row * wtiles;

    const long long gw = static_cast<long long>(blockIdx.x) * kPolyWarps + warp;
    long long row = gw / wtiles, tile = gw - row * wtiles;
    long long nrow = row, ntile = tile;                    // runs two tiles ahead
    auto advance = [&](long long& r, long long& t) {
        r += step_row;
        t += step_tile;
        if (t >= wtiles) {
            t -= wtiles;
            r += 1;
        }
    };
    auto first_sample = [&](long long t) {
        const long long m0 = q.base_m + t * WT_OUT;
        return (m0 / UP) * DOWN - HALO;
    };
    auto can_bulk = [&](long long a) { return q.in_vec_ok && a >= 0 && a + LEN <= q.n_in; };
    auto prefetch = [&](long long r, long long t, int s) { // lane 0 only
        if (r >= batch) return;
        const long long a = first_sample(t);
        if (!can_bulk(a)) return;
        mbar_arrive_expect_tx(bar0 + 8u * s, LEN * 4u);
        bulk_copy_g2s(smem_u32(mine + s * LEN), q.x + r * q.ld_x + a, LEN * 4u, bar0 + 8u * s);
    };

    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_mbar_init();
        prefetch(nrow, ntile, 0);
    }
    advance(nrow, ntile);
    if (lane == 0) prefetch(nrow, ntile, 1);
    advance(nrow, ntile);
    __syncwarp();

    uint32_t phases = 0;
    for (int it = 0; row < batch; ++it) {
        const int s = it & 1;
        const long long m0 = q.base_m + tile * WT_OUT;
        const long long a = first_sample(tile);
        float* const in = mine + s * LEN;
        if (can_bulk(a)) {
            mbar_wait(bar0 + 8u * s, (phases >> s) & 1u);
            phases ^= 1u << s;
        } else {                                           // edge tile: samples outside the row by extension mode
            const float* __restrict__ xr = q.x + row * q.ld_x;
            for (int t = lane; t < LEN; t += 32) in[t] = upfirdn_sample(xr, a + t, q.n_in, q.ext);
            __syncwarp();
        }

        float acc[R];
        poly_core2<UP, DOWN, G, KCP, Z>(acc, in + HALO + lane * SIN, taps, pairs);

        if (lane == 0) bulk_store_wait_read<0>();          // the previous tile's store has drained `out`
        __syncwarp();                                      // in[s] consumed by every lane, out free
        if (lane == 0) prefetch(nrow, ntile, s);           // refill in[s] two tiles ahead

        poly_store_acc<R>(out_s + lane * R, acc);

        float* __restrict__ yr = q.y + row * q.ld_y;
        if (q.out_vec_ok && m0 >= q.m_begin && m0 + WT_OUT <= q.m_end) {
            fence_proxy_async_smem();                      // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) bulk_copy_s2g(yr + (m0 - q.m_begin), smem_u32(out_s), WT_OUT * 4u);
        } else {
            __syncwarp();
            for (int t = lane; t < WT_OUT; t += 32) {
                const long long m = m0 + t;
                if (m >= q.m_begin && m < q.m_end) yr[m - q.m_begin] = out_s[t];
            }
            __syncwarp();                                  // out is rewritten by the next tile's stores
        }
        advance(row, tile);
        advance(nrow, ntile);
    }
    if (lane == 0) bulk_store_wait_read<0>();              // smem must outlive the last store's read
}

template <int UP, int DOWN, int G, int KCP, int Z, int NCH>
int launch_one(scir_b200_ctx* ctx, const PolyParams& q, const PolyTaps& taps, int nchunk, long long grid, long long batch)
{
    constexpr int R = UP * G, SIN = DOWN * G;
    const int halo = nchunk * KCP + ((Z > 0) ? 4 : 0);
    const size_t len = static_cast<size_t>(halo) + kPolyNT * SIN + 4;
    const size_t stream_bytes = (2 * len + static_cast<size_t>(kPolyNT) * R) * sizeof(float);
    const int d = ctx->device & 15;
    thread_local std::unique_ptr<PolyPairs> pairs_owner;    // freed when the thread exits
    if (!pairs_owner) pairs_owner.reset(new PolyPairs());
    PolyPairs* pairs = pairs_owner.get();
    // upfirdn_variant: 0 auto (warp-specialised, tap-reuse FFMA2 core) | 7 streaming + tap-reuse core | 8 streaming + one tap-pair
    // load per FFMA2 (the round-1 default) | 3 one tile per CTA (FFMA2) | 4 streaming, scalar FFMA | 5 tile, scalar FFMA |
    // 6 per-warp pipelines (FFMA2)
    const int v = static_cast<int>(ctx->opt.upfirdn_variant);
    int packed = (NCH == 1 && v != 4 && v != 5) ? 1 : 0;
    if (packed && (v == 0 || v == 7)) packed = 2;                     // 9: warp-specialised pipeline + the x-major core (A/B)
    if (packed == 2) PolyGeom<UP, DOWN, G, KCP, Z>::fill_pairs3(taps.c, pairs);
    else if (packed) PolyGeom<UP, DOWN, G, KCP, Z>::fill_pairs(taps.c, pairs);
    if constexpr (NCH == 1) {
        constexpr int WS_LEN = KCP + ((Z > 0) ? 4 : 0) + kPolyNT * SIN + 4;
        const int stages = (ctx->opt.upfirdn_ws_stages == 2) ? 2 : 3;      // 3 stages, 2 CTAs/SM (default) | 2 stages, 3 CTAs/SM
        const size_t ws_bytes = (static_cast<size_t>(stages) * WS_LEN + static_cast<size_t>(kPolyNT) * R) * sizeof(float);
        if ((v == 0 || v == 9) && ws_bytes <= static_cast<size_t>(ctx->max_smem_optin)) {
            auto kern = (v == 9) ? upfirdn_ws_kernel<UP, DOWN, G, KCP, Z, 3, 2>
                                 : (stages == 2) ? upfirdn_ws_kernel<UP, DOWN, G, KCP, Z, 2, 3> : upfirdn_ws_kernel<UP, DOWN, G, KCP, Z, 3, 3>;
            static thread_local size_t configured[16][3] = {};
            static thread_local int resident[16][3] = {};
            const int si = (v == 9) ? 2 : stages - 2;
            if (configured[d][si] < ws_bytes) {
                SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ws_bytes)),
                          "cudaFuncSetAttribute(upfirdn_ws_kernel)");
                configured[d][si] = ws_bytes;
                int nb = 0;
                SCIR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kWsThreads, ws_bytes),
                          "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
                resident[d][si] = std::max(nb, 1);
            }
            const long long g = std::min<long long>(grid, static_cast<long long>(ctx->sm_count) * resident[d][si]);
            kern<<<static_cast<unsigned>(g), kWsThreads, ws_bytes, ctx->stream>>>(q, taps, batch, *pairs);
            SCIR_CUDA(cudaGetLastError(), "upfirdn_ws_kernel launch");
            ctx->launches++;
            ctx->poly_launches++;
            return SCIR_B200_OK;
        }
        if (ctx->opt.upfirdn_variant == 6) {               // warp-pipelined A/B arm
            constexpr int LEN = KCP + ((Z > 0) ? 4 : 0) + 32 * SIN + 4, WT_OUT = 32 * R;
            const size_t bytes = static_cast<size_t>(kPolyWarps) * (2 * LEN + WT_OUT) * sizeof(float);
            auto kern = upfirdn_warp_kernel<UP, DOWN, G, KCP, Z>;
            static thread_local size_t configured[16] = {0};
            static thread_local int resident[16] = {0};
            if (configured[d] < bytes) {
                SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)),
                          "cudaFuncSetAttribute(upfirdn_warp_kernel)");
                configured[d] = bytes;
                int nb = 0;
                SCIR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kPolyWarps * 32, bytes),
                          "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
                resident[d] = std::max(nb, 1);
            }
            const long long wtiles = (q.m_end - q.base_m + WT_OUT - 1) / WT_OUT;
            const long long total_w = wtiles * batch;
            const long long g = std::min<long long>((total_w + kPolyWarps - 1) / kPolyWarps,
                                                    static_cast<long long>(ctx->sm_count) * resident[d]);
            kern<<<static_cast<unsigned>(g), kPolyWarps * 32, bytes, ctx->stream>>>(q, taps, batch, *pairs, wtiles);
            SCIR_CUDA(cudaGetLastError(), "upfirdn_warp_kernel launch");
            ctx->launches++;
            ctx->poly_launches++;
            return SCIR_B200_OK;
        }
    }
    if (ctx->opt.upfirdn_variant != 3 && ctx->opt.upfirdn_variant != 5 && stream_bytes <= static_cast<size_t>(ctx->max_smem_optin)) {
        // persistent grid: every SM holds as many CTAs as fit; each walks tiles with stride gridDim.x
        auto kern = upfirdn_stream_kernel<UP, DOWN, G, KCP, Z, NCH>;
        static thread_local size_t configured[16] = {0};
        static thread_local size_t occ_smem[16] = {0};
        static thread_local int resident[16] = {0};
        if (configured[d] < stream_bytes) {
            SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stream_bytes)),
                      "cudaFuncSetAttribute(upfirdn_stream_kernel)");
            configured[d] = stream_bytes;
        }
        if (resident[d] == 0 || occ_smem[d] != stream_bytes) {
            int nb = 0;
            SCIR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kPolyNT, stream_bytes),
                      "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            resident[d] = std::max(nb, 1);
            occ_smem[d] = stream_bytes;
        }
        const long long g = std::min<long long>(grid, static_cast<long long>(ctx->sm_count) * resident[d]);
        kern<<<static_cast<unsigned>(g), kPolyNT, stream_bytes, ctx->stream>>>(q, taps, batch, *pairs, packed);
        SCIR_CUDA(cudaGetLastError(), "upfirdn_stream_kernel launch");
        ctx->launches++;
        ctx->poly_launches++;
        return SCIR_B200_OK;
    }
    const size_t floats = std::max<size_t>(len, static_cast<size_t>(kPolyNT) * R);
    const size_t smem_bytes = floats * sizeof(float);
    if (smem_bytes > static_cast<size_t>(ctx->max_smem_optin))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "polyphase tile needs %zu B of shared memory", smem_bytes);
    auto kern = upfirdn_tile_kernel<UP, DOWN, G, KCP, Z, NCH>;
    static thread_local size_t configured[16] = {0};
    if (smem_bytes > 48 * 1024 && configured[d] < smem_bytes) {
        SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)),
                  "cudaFuncSetAttribute(upfirdn_tile_kernel)");
        configured[d] = smem_bytes;
    }
    kern<<<static_cast<unsigned>(grid), kPolyNT, smem_bytes, ctx->stream>>>(q, taps, *pairs, packed);
    SCIR_CUDA(cudaGetLastError(), "upfirdn_tile_kernel launch");
    ctx->launches++;
    ctx->poly_launches++;
    return SCIR_B200_OK;
}

// Z and NCH are runtime facts of the filter: dispatch onto the template grid.
template <int UP, int DOWN, int G, int KCP, int Z>
int launch_z(scir_b200_ctx* ctx, const PolyParams& q, const PolyTaps& taps, int nchunk, long long grid, long long batch)
{
    if (nchunk == 1) return launch_one<UP, DOWN, G, KCP, Z, 1>(ctx, q, taps, nchunk, grid, batch);
    return launch_one<UP, DOWN, G, KCP, Z, 0>(ctx, q, taps, nchunk, grid, batch);
}

template <int UP, int DOWN, int G, int KCP>
int launch_rate(scir_b200_ctx* ctx, const PolyParams& q, const PolyTaps& taps, int nchunk, int z, long long grid, long long batch)
{
    if constexpr (UP >= 2) {
        if (z == 1) return launch_z<UP, DOWN, G, KCP, 1>(ctx, q, taps, nchunk, grid, batch);
    }
    if constexpr (UP >= 3) {
        if (z == 2) return launch_z<UP, DOWN, G, KCP, 2>(ctx, q, taps, nchunk, grid, batch);
    }
    if constexpr (UP >= 4) {
        if (z == 3) return launch_z<UP, DOWN, G, KCP, 3>(ctx, q, taps, nchunk, grid, batch);
    }
    return launch_z<UP, DOWN, G, KCP, 0>(ctx, q, taps, nchunk, grid, batch);
}

struct RateGeom {
    int up, down, g;
};
// G per rate: (down*G) % 4 == 0 and (down*G/4) odd; R = up*G outputs per thread.
constexpr RateGeom kRates[] = {
    {3, 2, 10}, {2, 3, 12}, {2, 1, 12}, {1, 2, 10}, {3, 1, 4}, {1, 3, 12}, {4, 1, 4}, {1, 4, 5}, {4, 3, 4}, {3, 4, 5},
};

}  // namespace

int launch_upfirdn_poly(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                        const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y, int64_t ld_y,
                        int64_t m_begin, int64_t m_count, ExtSpec ext, bool* handled)
{
    *handled = false;
    constexpr int KCP = 32;
    int g = 0;
    for (const auto& r : kRates)
        if (r.up == up && r.down == down) g = r.g;
    if (g == 0) return SCIR_B200_OK;                                  // rate not in the template grid

    // structural zeros at both ends of the filter (resample_poly's pre/post padding)
    int64_t lo = 0, hi = len_h;
    while (hi > 1 && h[hi - 1] == 0.f) --hi;
    while (lo < hi - 1 && h[lo] == 0.f) ++lo;
    const int z = (lo < up) ? static_cast<int>(lo) : 0;
    int64_t kp = 1;                                                   // taps per phase after skipping
    for (int64_t t = 0; t < up; ++t) {
        if (hi - 1 - t < 0) continue;
        const int64_t i_hi = (hi - 1 - t) / up;
        const int64_t o_t = (t < z) ? 1 : 0;
        kp = std::max<int64_t>(kp, i_hi - o_t + 1);
    }
    const int64_t nchunk = (kp + KCP - 1) / KCP;
    if ((nchunk * KCP + 2) * up > kPolyMaxH) return SCIR_B200_OK;     // filter too long for the parameter block

    thread_local std::unique_ptr<PolyTaps> tl;              // freed when the thread exits
    if (!tl) tl.reset(new PolyTaps());
    for (int i = 0; i < kPolyMaxH; ++i) tl->c[i] = (i < hi) ? h[i] : 0.f;

    PolyParams q{};
    q.x = d_x; q.y = d_y; q.ld_x = ld_x; q.ld_y = ld_y; q.n_in = n_in;
    q.m_begin = m_begin; q.m_end = m_begin + m_count;
    q.ext = ext;
    const int64_t step = 4 * up;
    q.base_m = (m_begin / step) * step;
    const int64_t tile_out = static_cast<int64_t>(kPolyNT) * up * g;
    q.ntiles = (q.m_end - q.base_m + tile_out - 1) / tile_out;
    q.nchunk = static_cast<int>(nchunk);
    const bool generic_io = (ctx->opt.upfirdn_variant == 2);
    q.in_vec_ok = !generic_io && aligned16(d_x) && (ld_x % 4 == 0);
    q.out_vec_ok = !generic_io && aligned16(d_y) && (ld_y % 4 == 0) && ((m_begin - q.base_m) % 4 == 0);
    const long long grid = q.ntiles * batch;
    if (grid > 0x7fffffffLL) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));

    int rc = SCIR_B200_OK;
#define SCIR_RATE(U, D, GG) \
    if (up == U && down == D) { rc = launch_rate<U, D, GG, KCP>(ctx, q, *tl, q.nchunk, z, grid, batch); *handled = (rc == SCIR_B200_OK); return rc; }
    SCIR_RATE(3, 2, 10)
    SCIR_RATE(2, 3, 12)
    SCIR_RATE(2, 1, 12)
    SCIR_RATE(1, 2, 10)
    SCIR_RATE(3, 1, 4)
    SCIR_RATE(1, 3, 12)
    SCIR_RATE(4, 1, 4)
    SCIR_RATE(1, 4, 5)
    SCIR_RATE(4, 3, 4)
    SCIR_RATE(3, 4, 5)
#undef SCIR_RATE
    return rc;
}

}  // namespace scir_b200
