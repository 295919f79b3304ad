// fir_direct.cu -- FP32 direct-form batched FIR for sm_100a.
//
// Replaces the reference's one-thread-per-output PTX entry (crates/scir-gpu/src/lib.rs:727-811,
// which does 2 global loads per MAC and does not even assemble) with a tile kernel designed around
// the B200 issue model: the FP32 pipe retires one FFMA warp-instruction per cycle per SM
// sub-partition, which is also the scheduler's whole issue budget, so every non-FFMA instruction
// is lost FP32 throughput.  The design therefore drives the FFMA share of the instruction stream
// towards 100 %:
//
//   * taps are LAUNCH PARAMETERS (constant bank 0).  ptxas keeps the current chunk of KC taps in
//     UNIFORM registers (LDCU) and emits `FFMA Racc, Rx.reuse, URtap, Racc`: two register
//     operands, no register-bank conflict, no per-thread tap loads, no global constant state.
//   * each thread owns R consecutive outputs (R accumulators) and walks its input window
//     x-major: one LDS.128 brings 4 samples, each sample feeds up to R FFMAs.  FFMA:LDS = R*KC/(R+KC)*4.
//   * R = 20: the per-thread window stride is 80 B, so the 8 lanes of an LDS.128 phase hit
//     8 distinct 16-B bank groups (R/4 odd) -- conflict-free without padding or swizzle, which is
//     what lets the tile be brought in by ONE 1-D TMA bulk copy (cp.async.bulk, dense layout).
//   * the tile plus its K-1 sample halo arrives in shared memory by cp.async.bulk + mbarrier
//     (issued by one thread: zero issue slots on the compute warps); results go back through
//     shared memory and ONE cp.async.bulk store (fully coalesced, again one instruction).
//   * several CTAs are resident per SM (40 registers/thread), so one CTA's copy-in / copy-out
//     overlaps its neighbours' FFMA streams; no tracing compiler, no software pipeline needed.
//   * tap counts of any size run the same code: a runtime loop over chunks of KC taps
//     (`LDCU UR, c[0x0][UR+imm]`), halo = nchunk*KC samples.
//   * for K <= 256 the inner product is issued as packed FFMA2 (fir_core2): two outputs per issue slot, the
//     sample broadcast, two consecutive taps from a uniform-register pair -- half the issue slots for the same
//     FMA pipe rate, so loads and loop control stop costing FP32 throughput (config 2: 2.75 -> 2.25 ms).
//
// Since the tcgen05 Toeplitz kernel (fir_toeplitz.cu) takes every large launch, this family serves small and
// mid-size launches (api.cu: prefer_toeplitz) and the A/B arms.
//
// The same kernel serves lfilter / filtfilt: direction (causal / anticausal), boundary rule
// (zero state or held end value = SciPy's lfilter_zi steady state) and odd/even/constant signal
// extension are properties of the tile LOADER (edge tiles only); interior tiles always take the
// bulk-copy path.  See DESIGN.md section 4.
#include "common.cuh"

#include <memory>
#include "ptx.cuh"

#include <algorithm>

namespace scir_b200 {

constexpr int kPackedMaxK = 256;      // instantiations up to this many taps carry the shifted copy and use FFMA2

template <int MAXK>
struct TapsParam {
    float c[MAXK];
    // c1[i] = c[i+1]: makes the tap pair (c[d], c[d+1]) an aligned 64-bit constant load for odd d too (fir_core2).
    // Only the short-filter instantiations carry it (the 7936-tap one would not fit the 32 KB parameter space).
    float c1[(MAXK <= kPackedMaxK) ? MAXK : 2];
};

struct FirTileParams {
    FirPass p;
    long long base0;      // virtual index of the first output of tile 0 ((base0 + in_off) % 4 == 0)
    long long ntiles;     // tiles per row
    int nchunk;           // tap chunks (halo = nchunk * KC)
    int k;                // true tap count (non-finite redo)
    int packed;           // 1: FFMA2 core (fir_core2), 0: scalar FFMA core (A/B: ctx option "ffma2" = 0)
    int in_vec_ok;        // input rows are 16-B aligned => interior tiles may use the bulk copy
    int out_vec_ok;       // same for the output side
};

// ---- virtual input sequence (edge tiles only) ------------------------------------------------------
__device__ __forceinline__ float vload(const FirPass& p, const float* __restrict__ xr, long long i)
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
        if (p.ext_mode == EXT_ODD) return 2.f * xr[0] - xr[-u];      // scipy _arraytools.py:57-107
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

// ---- the FFMA core: R outputs per thread, x-major over the window, taps from uniform registers ------
// causal:     wbase points at v[t0] (smem index DP + tid*R); chunk c uses samples t0 + r - d, d in [c*KC,(c+1)*KC)
// anticausal: wbase points at v[t0] (smem index tid*R);      chunk c uses samples t0 + r + d
template <int KC, int R, int DIR, int MAXK>
__device__ __forceinline__ void fir_core(float (&acc)[R], const float* wbase, int nchunk, const TapsParam<MAXK>& taps)
{
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
        const float4* w = reinterpret_cast<const float4*>((DIR > 0) ? (wbase - (c + 1) * KC) : (wbase + c * KC));
        const float* tc = taps.c + c * KC;
#pragma unroll
        for (int v4 = 0; v4 < (KC + R) / 4; ++v4) {
            const float4 v = w[v4];
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = 4 * v4 + e;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    // causal: sample index (rel. to t0) = s - KC  => local delay = r - (s - KC)
                    // anticausal: sample index = s                => local delay = s - r
                    const int dl = (DIR > 0) ? (r + KC - s) : (s - r);
                    if (dl >= 0 && dl < KC) acc[r] = fmaf(tc[dl], xv[e], acc[r]);
                }
            }
        }
    }
}

// ---- packed form of the same core (short-filter instantiations) ----------------------------------------------
// sm_100's FFMA2 does two FMAs per issue slot: `FFMA2 Racc.F32x2, Rx.F32, URtaps.F32x2, Racc.F32x2` -- the sample
// broadcast, two consecutive taps from a uniform-register pair, two consecutive outputs.  At the same FMA pipe
// rate only half the issue slots are arithmetic, so LDS / LDCU / loop control stop costing FP32 throughput
// (microbenchmark: 74.0 vs 67.7 TFLOP/s for the scalar stream).  Outputs (2p, 2p+1) of a thread share every
// sample x[s]; their taps are (c[d], c[d+1]) (causal) or (c[d], c[d-1]) (anticausal): one aligned LDCU.64 from
// `c` (even first index) or from the shifted copy `c1` (odd).  Where only one of the two outputs has a tap in
// this chunk the other half multiplies by zero (a few percent of the instructions, at the chunk edges only).
__device__ __forceinline__ void fma2_x(unsigned long long& acc, float x, float t0, float t1)
{
    unsigned long long xx, tt;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(tt) : "f"(t0), "f"(t1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(tt));
}

template <int KC, int R, int DIR, int MAXK>
__device__ __forceinline__ void fir_core2(float (&acc)[R], const float* wbase, int nchunk, const TapsParam<MAXK>& taps)
{
    static_assert(R % 2 == 0 && KC % 2 == 0 && MAXK <= kPackedMaxK, "packed core geometry");
    unsigned long long acc2[R / 2];
#pragma unroll
    for (int p = 0; p < R / 2; ++p) acc2[p] = 0ull;
    for (int c = 0; c < nchunk; ++c) {
        const float4* w = reinterpret_cast<const float4*>((DIR > 0) ? (wbase - (c + 1) * KC) : (wbase + c * KC));
        const float* tc = taps.c + c * KC;
        const float* tc1 = taps.c1 + c * KC;
#pragma unroll
        for (int v4 = 0; v4 < (KC + R) / 4; ++v4) {
            const float4 v = w[v4];
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = 4 * v4 + e;
#pragma unroll
                for (int p = 0; p < R / 2; ++p) {
                    // chunk-local delays of outputs 2p and 2p+1 for this sample
                    const int d0 = (DIR > 0) ? (2 * p + KC - s) : (s - 2 * p);
                    const int d1 = (DIR > 0) ? (d0 + 1) : (d0 - 1);
                    const bool ok0 = d0 >= 0 && d0 < KC, ok1 = d1 >= 0 && d1 < KC;
                    if (ok0 && ok1) {
                        const int lo = (DIR > 0) ? d0 : d1;                    // the pair's lower tap index
                        const float2 t = (lo % 2 == 0) ? *reinterpret_cast<const float2*>(tc + lo)
                                                       : *reinterpret_cast<const float2*>(tc1 + lo - 1);
                        if (DIR > 0) fma2_x(acc2[p], xv[e], t.x, t.y);
                        else fma2_x(acc2[p], xv[e], t.y, t.x);
                    } else if (ok0) {
                        fma2_x(acc2[p], xv[e], tc[d0], 0.f);
                    } else if (ok1) {
                        fma2_x(acc2[p], xv[e], 0.f, tc[d1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < R / 2; ++p) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * p]), "=f"(acc[2 * p + 1]) : "l"(acc2[p]));
}

// dispatch: packed core where the instantiation carries the shifted tap copy
template <int KC, int R, int DIR, int MAXK>
__device__ __forceinline__ void fir_core_any(float (&acc)[R], const float* wbase, int nchunk, const TapsParam<MAXK>& taps, int packed)
{
    if constexpr (MAXK <= kPackedMaxK) {
        if (packed) {
            fir_core2<KC, R, DIR, MAXK>(acc, wbase, nchunk, taps);
            return;
        }
    }
    fir_core<KC, R, DIR, MAXK>(acc, wbase, nchunk, taps);
}

// Non-finite data: the tap array is zero-padded to whole chunks, and 0 * Inf = NaN would leak a stray Inf / NaN
// into up to KC-1 outputs whose true window does not contain it.  A thread whose outputs came out non-finite
// (one add per output to find out) redoes them from the staged window with exactly k taps, the way the
// reference loop does (lib.rs:1138-1150); with finite data the branch is never taken.
template <int R>
__device__ __forceinline__ bool any_nonfinite(const float (&acc)[R])
{
    // Inf and NaN survive a sum (Inf - Inf = NaN); a finite sum that overflows only costs a harmless redo
    float s = acc[0];
#pragma unroll
    for (int r = 1; r < R; ++r) s += acc[r];
    return (__float_as_uint(s) & 0x7f800000u) == 0x7f800000u;
}

// w0: the thread's first output position in the staged window (smem index DP + tid*R causal, tid*R anticausal)
template <int R, int DIR, int MAXK>
__device__ __forceinline__ void exact_redo(float (&acc)[R], const float* w0, int k, const TapsParam<MAXK>& taps)
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float a = 0.f;
        for (int d = 0; d < k; ++d) a = fmaf(taps.c[d], (DIR > 0) ? w0[r - d] : w0[r + d], a);
        acc[r] = a;
    }
}

// ---- one tile per CTA (variant 3; the round-1 first cut, kept for A/B profiling) ---------------------
// KC taps per chunk, R outputs per thread, NT threads, DIR +1 causal / -1 anticausal.
template <int KC, int R, int NT, int DIR, int MAXK>
__global__ void __launch_bounds__(NT, (NT <= 128) ? 8 : 4)
fir_tile_kernel(const __grid_constant__ FirTileParams q, const __grid_constant__ TapsParam<MAXK> taps)
{
    static_assert(KC % 4 == 0 && R % 4 == 0, "vector width");
    static_assert((R / 4) % 2 == 1, "R/4 must be odd: conflict-free LDS.128 on a dense tile");
    constexpr int TILE = NT * R;
    extern __shared__ __align__(128) float smem[];     // [DP + TILE]
    __shared__ __align__(8) unsigned long long mbar;

    const FirPass& p = q.p;
    const int tid = threadIdx.x;
    const int DP = q.nchunk * KC;
    const int len = DP + TILE;
    const long long bid = blockIdx.x;
    const long long row = bid / q.ntiles;
    const long long tile = bid - row * q.ntiles;
    const long long i0 = q.base0 + tile * TILE;          // first output of this tile (virtual index)
    const long long a = (DIR > 0) ? (i0 - DP) : i0;       // first sample staged in smem
    const float* __restrict__ xr = p.x + row * p.ld_x;
    float* __restrict__ yr = p.y + row * p.ld_y;

    // ---- stage the tile + halo ---------------------------------------------------------------------
    bool bulk_in = q.in_vec_ok && a >= 0 && a + len <= p.n_v;
    if (p.ext_mode != EXT_NONE) bulk_in = bulk_in && (a + p.in_off >= 0) && (a + p.in_off + len <= p.n_x);
    if (bulk_in) {
        const uint32_t bar = smem_u32(&mbar);
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(bar, static_cast<uint32_t>(len) * 4u);
            bulk_copy_g2s(smem_u32(smem), xr + a + p.in_off, static_cast<uint32_t>(len) * 4u, bar);
        }
        __syncthreads();                                   // barrier init visible to the waiters
        mbar_wait(bar, 0);
    } else {
        for (int s = tid; s < len; s += NT) smem[s] = vload(p, xr, a + s);
        __syncthreads();
    }

    float acc[R];
    fir_core_any<KC, R, DIR, MAXK>(acc, (DIR > 0) ? (smem + DP + tid * R) : (smem + tid * R), q.nchunk, taps, q.packed);

    // ---- results back through shared memory, then one bulk store --------------------------------------
    if (any_nonfinite<R>(acc)) exact_redo<R, DIR, MAXK>(acc, (DIR > 0) ? (smem + DP + tid * R) : (smem + tid * R), q.k, taps);
    __syncthreads();                                       // every warp is done reading the tile
    float4* so = reinterpret_cast<float4*>(smem + tid * R);
#pragma unroll
    for (int r = 0; r < R; r += 4) so[r / 4] = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);

    const bool bulk_out = q.out_vec_ok && i0 >= p.out_begin && i0 + TILE <= p.out_end;
    if (bulk_out) {
        fence_proxy_async_smem();                          // generic-proxy writes -> async proxy
        __syncthreads();
        if (tid == 0) {
            bulk_copy_s2g(yr + i0 + p.out_off, smem_u32(smem), TILE * 4u);
            bulk_store_wait_read<0>();                        // smem must outlive the read
        }
    } else {
        __syncthreads();
        for (int s = tid; s < TILE; s += NT) {
            const long long i = i0 + s;
            if (i >= p.out_begin && i < p.out_end) yr[i + p.out_off] = smem[s];
        }
    }
}

// ---- the streaming kernel (default): persistent CTAs, double-buffered TMA bulk loads ------------------
// Each CTA walks tiles w = blockIdx.x, +gridDim.x, ... of the (row, tile) grid.  Two input stages are
// filled by cp.async.bulk two tiles ahead (full[] mbarriers, one elected thread), results leave
// through a third buffer and a bulk store that drains while the next tile is being computed.  The
// per-tile cost outside the FFMA stream is two CTA barriers, five STS.128 and a dozen scalar
// instructions; index set-up, barrier init and the first DRAM round trip are paid once per CTA
// instead of once per tile.
struct TileCoord {
    long long row, tile;
};

template <int KC, int R, int NT, int DIR, int MAXK>
__global__ void __launch_bounds__(NT, 3)
fir_stream_kernel(const __grid_constant__ FirTileParams q, const __grid_constant__ TapsParam<MAXK> taps)
{
    static_assert(KC % 4 == 0 && R % 4 == 0, "vector width");
    static_assert((R / 4) % 2 == 1, "R/4 must be odd: conflict-free LDS.128 on a dense tile");
    constexpr int TILE = NT * R;
    extern __shared__ __align__(128) float smem[];     // in[0] | in[1] | out
    __shared__ __align__(8) unsigned long long full[2];

    const FirPass& p = q.p;
    const int tid = threadIdx.x;
    const int DP = q.nchunk * KC;
    const int len = DP + TILE;
    float* const out_s = smem + 2 * len;
    const uint32_t bar0 = smem_u32(&full[0]);              // full[s] lives at bar0 + 8*s
    const long long G = gridDim.x;
    const long long step_row = G / q.ntiles, step_tile = G - step_row * q.ntiles;

    auto advance = [&](TileCoord& t) {
        t.row += step_row;
        t.tile += step_tile;
        if (t.tile >= q.ntiles) {
            t.tile -= q.ntiles;
            t.row += 1;
        }
    };
    auto first_sample = [&](const TileCoord& t) {          // virtual index of the first staged sample
        const long long i0 = q.base0 + t.tile * TILE;
        return (DIR > 0) ? (i0 - DP) : i0;
    };
    auto can_bulk = [&](long long a) {
        bool ok = q.in_vec_ok && a >= 0 && a + len <= p.n_v;
        if (p.ext_mode != EXT_NONE) ok = ok && (a + p.in_off >= 0) && (a + p.in_off + len <= p.n_x);
        return ok;
    };
    auto prefetch = [&](const TileCoord& t, int s) {       // thread 0 only
        if (t.row >= p.batch) return;
        const long long a = first_sample(t);
        if (!can_bulk(a)) return;
        mbar_arrive_expect_tx(bar0 + 8u * s, static_cast<uint32_t>(len) * 4u);
        bulk_copy_g2s(smem_u32(smem + s * len), p.x + t.row * p.ld_x + a + p.in_off, static_cast<uint32_t>(len) * 4u,
                      bar0 + 8u * s);
    };

    TileCoord cur;
    cur.row = blockIdx.x / q.ntiles;
    cur.tile = blockIdx.x - cur.row * q.ntiles;
    TileCoord nxt = cur;                                   // runs two tiles ahead of cur
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_mbar_init();
        prefetch(nxt, 0);
        advance(nxt);
        prefetch(nxt, 1);
        advance(nxt);
    } else {
        advance(nxt);
        advance(nxt);
    }
    __syncthreads();                                       // barrier init visible to the waiters

    uint32_t phases = 0;                                   // bit s = parity to wait for on full[s]
    for (int it = 0; cur.row < p.batch; ++it, advance(cur), advance(nxt)) {
        const int s = it & 1;
        const long long i0 = q.base0 + cur.tile * TILE;
        const long long a = (DIR > 0) ? (i0 - DP) : i0;
        const float* __restrict__ xr = p.x + cur.row * p.ld_x;
        float* __restrict__ yr = p.y + cur.row * p.ld_y;
        float* const in = smem + s * len;

        if (can_bulk(a)) {
            mbar_wait(bar0 + 8u * s, (phases >> s) & 1u);
            phases ^= 1u << s;
        } else {                                           // edge tile: zero / held / extended samples
            for (int t = tid; t < len; t += NT) in[t] = vload(p, xr, a + t);
            __syncthreads();
        }

        float acc[R];
        fir_core<KC, R, DIR, MAXK>(acc, (DIR > 0) ? (in + DP + tid * R) : (in + tid * R), q.nchunk, taps);   // scalar A/B arm

        if (tid == 0) bulk_store_wait_read<0>();           // the previous tile's store has drained `out`
        if (any_nonfinite<R>(acc)) exact_redo<R, DIR, MAXK>(acc, (DIR > 0) ? (in + DP + tid * R) : (in + tid * R), q.k, taps);
        __syncthreads();                                   // A: in[s] consumed by every warp, out free
        if (tid == 0) prefetch(nxt, s);                    // refill in[s] two tiles ahead

        float4* so = reinterpret_cast<float4*>(out_s + tid * R);
#pragma unroll
        for (int r = 0; r < R; r += 4) so[r / 4] = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);

        const bool bulk_out = q.out_vec_ok && i0 >= p.out_begin && i0 + TILE <= p.out_end;
        if (bulk_out) {
            fence_proxy_async_smem();                      // generic-proxy writes -> async proxy
            __syncthreads();                               // B
            if (tid == 0) bulk_copy_s2g(yr + i0 + p.out_off, smem_u32(out_s), TILE * 4u);
        } else {
            __syncthreads();
            for (int t = tid; t < TILE; t += NT) {
                const long long i = i0 + t;
                if (i >= p.out_begin && i < p.out_end) yr[i + p.out_off] = out_s[t];
            }
        }
    }
    if (tid == 0) bulk_store_wait_read<0>();               // smem must outlive the last store's read
}

#ifndef SCIR_FIR_DIRECT_REV   // the direction-independent pieces live in the forward translation unit only
// ---- naive kernel: one thread per output, taps from global memory ------------------------------------
// An independent implementation of the same FirPass (what the reference's PTX entry was meant to
// do, lib.rs:744-809).  Kept for A/B profiling (`variant=2`) and as a cross-check in the tests.
__global__ void fir_naive_kernel(const FirPass p, const float* __restrict__ c, int k)
{
    const long long per_row = p.out_end - p.out_begin;
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= per_row * p.batch) return;
    const long long row = gid / per_row;
    const long long i = p.out_begin + (gid - row * per_row);
    const float* xr = p.x + row * p.ld_x;
    float acc = 0.f;
    for (int d = 0; d < k; ++d) acc = fmaf(c[d], vload(p, xr, p.dir > 0 ? i - d : i + d), acc);
    p.y[row * p.ld_y + i + p.out_off] = acc;
}

// ---- lfilter streaming state -----------------------------------------------------------------------------
// y[b, j] += zi[b, j] for j < min(k-1, n)   (scipy/signal/_signaltools.py:2230-2232)
__global__ void add_zi_kernel(float* y, long long ld_y, const float* __restrict__ zi, long long batch,
                              long long n, int km1)
{
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= batch * km1) return;
    const long long row = gid / km1;
    const int j = static_cast<int>(gid - row * km1);
    if (j < n) y[row * ld_y + j] += zi[row * km1 + j];
}

// zf[b, j] = full[n + j], full = convolve(b, x) (+ zi on its first k-1 entries)  (:2238-2240)
__global__ void compute_zf_kernel(const float* __restrict__ b, int k, const float* __restrict__ x,
                                  long long ld_x, const float* __restrict__ zi, float* zf,
                                  long long batch, long long n)
{
    const int km1 = k - 1;
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= batch * km1) return;
    const long long row = gid / km1;
    const int j = static_cast<int>(gid - row * km1);
    const float* xr = x + row * ld_x;
    float acc = 0.f;
    for (int d = j + 1; d < k; ++d) {          // full[n+j] = sum_d b[d] * x[n + j - d], index < n
        const long long xi = n + j - d;
        if (xi >= 0) acc = fmaf(b[d], xr[xi], acc);
    }
    if (zi != nullptr && n + j < km1) acc += zi[row * km1 + n + j];
    zf[row * km1 + j] = acc;
}

#endif  // !SCIR_FIR_DIRECT_REV

// ---- host side ---------------------------------------------------------------------------------------------
namespace {

constexpr int kR = 20;
constexpr int kNT = 256;
constexpr int kTile = kR * kNT;
constexpr int kSmallK = 256;
constexpr int kBigK = SCIR_B200_MAX_TAPS;

template <int KC, int DIR, int MAXK>
int launch_tile(scir_b200_ctx* ctx, const FirTileParams& q, const float* c, int64_t k, long long total)
{
    // zero-padded host staging; the launch copies it into the parameter buffer synchronously
    thread_local std::unique_ptr<TapsParam<MAXK>> tl;       // freed when the thread exits
    if (!tl) tl.reset(new TapsParam<MAXK>());
    for (int i = 0; i < MAXK; ++i) tl->c[i] = (i < k) ? c[i] : 0.f;
    if (MAXK <= kPackedMaxK)
        for (int i = 0; i < MAXK; ++i) tl->c1[i] = (i + 1 < k) ? c[i + 1] : 0.f;
    const size_t len = static_cast<size_t>(q.nchunk) * KC + kTile;
    // One tile per CTA (4 CTAs/SM overlap each other's copies) is the default: with the packed FFMA2 core it beats the
    // persistent CTA-streaming kernel at every size (profiles/r01_dispatch_sweep.txt: 10-20 % for K <= 128), whose
    // two CTA barriers per tile are no longer hidden behind FFMA issue.  Streaming stays as the scalar-core default
    // for short filters (ffma2=0) and as variant 4; it carries the scalar core only.
    const bool packed = (MAXK <= kPackedMaxK && ctx->opt.ffma2 != 0);
    const bool stream = (ctx->opt.variant == 4) || (ctx->opt.variant != 3 && q.nchunk <= 2 && !packed);
    const size_t smem_bytes = (stream ? (2 * len + kTile) : len) * sizeof(float);
    if (smem_bytes > static_cast<size_t>(ctx->max_smem_optin))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "tile needs %zu B of shared memory", smem_bytes);
    auto kern = stream ? fir_stream_kernel<KC, kR, kNT, DIR, MAXK> : fir_tile_kernel<KC, kR, kNT, DIR, MAXK>;
    // per (thread, device, kernel flavour): opt-in shared memory and the resident-CTA count
    static thread_local size_t configured[16][2] = {};
    static thread_local int resident[16][2] = {};
    const int d = ctx->device & 15, f = stream ? 1 : 0;
    if (configured[d][f] < smem_bytes || resident[d][f] == 0) {
        SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(std::max<size_t>(smem_bytes, configured[d][f]))),
                  "cudaFuncSetAttribute(fir kernel)");
        configured[d][f] = std::max<size_t>(smem_bytes, configured[d][f]);
        resident[d][f] = -1;
    }
    long long grid = total;
    if (stream) {
        // persistent grid: every SM holds as many CTAs as fit; each walks tiles with stride gridDim.x
        static thread_local size_t occ_smem[16] = {};
        if (resident[d][f] <= 0 || occ_smem[d] != smem_bytes) {
            int nb = 0;
            SCIR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kNT, smem_bytes),
                      "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            resident[d][f] = std::max(nb, 1);
            occ_smem[d] = smem_bytes;
        }
        grid = std::min<long long>(total, static_cast<long long>(ctx->sm_count) * resident[d][f]);
    }
    FirTileParams qq = q;
    qq.k = static_cast<int>(k);
    qq.packed = (MAXK <= kPackedMaxK && ctx->opt.ffma2 != 0) ? 1 : 0;
    kern<<<static_cast<unsigned>(grid), kNT, smem_bytes, ctx->stream>>>(qq, *tl);
    SCIR_CUDA(cudaGetLastError(), "fir kernel launch");
    ctx->launches++;
    return SCIR_B200_OK;
}

}  // namespace

// The causal and the anticausal instantiations compile in separate translation units (fir_direct_rev.cu includes
// this file with SCIR_FIR_DIRECT_REV defined): the fully unrolled cores make this the slowest file to build.
#ifdef SCIR_FIR_DIRECT_REV
int launch_fir_tiles_rev(scir_b200_ctx* ctx, const FirTileParams& q, const float* c, int64_t k, long long grid, int kc, bool small)
{
    if (kc == 32) return launch_tile<32, -1, kSmallK>(ctx, q, c, k, grid);
    if (small) return launch_tile<64, -1, kSmallK>(ctx, q, c, k, grid);
    return launch_tile<64, -1, kBigK>(ctx, q, c, k, grid);
}
#else
int launch_fir_tiles_rev(scir_b200_ctx* ctx, const FirTileParams& q, const float* c, int64_t k, long long grid, int kc, bool small);
static int launch_fir_tiles_fwd(scir_b200_ctx* ctx, const FirTileParams& q, const float* c, int64_t k, long long grid, int kc, bool small)
{
    if (kc == 32) return launch_tile<32, +1, kSmallK>(ctx, q, c, k, grid);
    if (small) return launch_tile<64, +1, kSmallK>(ctx, q, c, k, grid);
    return launch_tile<64, +1, kBigK>(ctx, q, c, k, grid);
}

int launch_fir_pass(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k)
{
    if (k < 1 || k > SCIR_B200_MAX_TAPS)
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "tap count %lld outside [1, %d]", (long long)k,
                         SCIR_B200_MAX_TAPS);
    if (pass.batch <= 0 || pass.out_end <= pass.out_begin) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));

    if (ctx->opt.variant == 2) {                      // naive A/B kernel: taps via a device buffer
        float* d_c = nullptr;
        SCIR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_c), static_cast<size_t>(k) * 4, ctx->stream),
                  "cudaMallocAsync(taps)");
        SCIR_CUDA(cudaMemcpyAsync(d_c, c, static_cast<size_t>(k) * 4, cudaMemcpyHostToDevice, ctx->stream),
                  "cudaMemcpyAsync(taps)");
        const long long total = (pass.out_end - pass.out_begin) * pass.batch;
        const long long blocks = (total + 255) / 256;
        if (blocks > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "grid too large");
        fir_naive_kernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(pass, d_c, static_cast<int>(k));
        SCIR_CUDA(cudaGetLastError(), "fir_naive_kernel launch");
        ctx->launches++;
        SCIR_CUDA(cudaFreeAsync(d_c, ctx->stream), "cudaFreeAsync(taps)");
        return SCIR_B200_OK;
    }

    // variant: 0 auto | 1 generic (non-bulk) IO | 2 naive | 3 one-tile-per-CTA | 4 CTA-streaming
    FirTileParams q;
    q.p = pass;
    const int KC = (k <= 32) ? 32 : 64;
    q.nchunk = static_cast<int>((k + KC - 1) / KC);
    const long long DP = static_cast<long long>(q.nchunk) * KC;
    const bool force_generic = (ctx->opt.variant == 1);
    q.in_vec_ok = !force_generic && aligned16(pass.x) && (pass.ld_x % 4 == 0);
    // tile origin: largest value <= out_begin whose input address is 16-B aligned
    const long long mis = (((pass.out_begin + pass.in_off) % 4) + 4) % 4;
    q.base0 = pass.out_begin - mis;
    q.out_vec_ok = !force_generic && aligned16(pass.y) && (pass.ld_y % 4 == 0) &&
                   ((((q.base0 + pass.out_off) % 4) + 4) % 4 == 0);
    q.ntiles = (pass.out_end - q.base0 + kTile - 1) / kTile;
    const long long grid = q.ntiles * pass.batch;
    if (grid > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "grid too large (%lld tiles)", grid);

    const bool small = (DP <= kSmallK);
    return (pass.dir > 0) ? launch_fir_tiles_fwd(ctx, q, c, k, grid, KC, small) : launch_fir_tiles_rev(ctx, q, c, k, grid, KC, small);
}

int launch_add_zi(scir_b200_ctx* ctx, float* d_y, int64_t ld_y, const float* d_zi, int64_t batch,
                  int64_t n, int64_t k)
{
    if (k <= 1 || batch <= 0 || n <= 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    const long long total = batch * (k - 1);
    add_zi_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(
        d_y, ld_y, d_zi, batch, n, static_cast<int>(k - 1));
    SCIR_CUDA(cudaGetLastError(), "add_zi_kernel launch");
    ctx->launches++;
    return SCIR_B200_OK;
}

int launch_compute_zf(scir_b200_ctx* ctx, const float* b, int64_t k, const float* d_x, int64_t ld_x,
                      const float* d_zi, float* d_zf, int64_t batch, int64_t n)
{
    if (k <= 1 || batch <= 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    float* d_b = nullptr;
    SCIR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_b), static_cast<size_t>(k) * 4, ctx->stream),
              "cudaMallocAsync(b)");
    SCIR_CUDA(cudaMemcpyAsync(d_b, b, static_cast<size_t>(k) * 4, cudaMemcpyHostToDevice, ctx->stream),
              "cudaMemcpyAsync(b)");
    const long long total = batch * (k - 1);
    compute_zf_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(
        d_b, static_cast<int>(k), d_x, ld_x, d_zi, d_zf, batch, n);
    SCIR_CUDA(cudaGetLastError(), "compute_zf_kernel launch");
    ctx->launches++;
    SCIR_CUDA(cudaFreeAsync(d_b, ctx->stream), "cudaFreeAsync(b)");
    return SCIR_B200_OK;
}

#endif  // SCIR_FIR_DIRECT_REV

}  // namespace scir_b200
