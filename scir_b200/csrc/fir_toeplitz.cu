// fir_toeplitz.cu -- block-Toeplitz FIR on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same FirPass contract as fir_direct.cu (reference hot path crates/scir-gpu/src/lib.rs:1134-1152 and
// its lfilter / filtfilt variants), but the sliding window is contracted on the tensor pipe:
//
//   a row is cut into blocks of B = 128 samples, X[s, j] = u[128 j + s]; then
//       Y[:, j] = sum_{p=0..P} T_p X[:, j-p],     T_p[r, s] = c[128 p + r - s]   (Toeplitz blocks)
//   i.e. per p one MMA  D[128 x N] += A[128 x 128] * B[128 x N]  with D in TMEM (FP32).
//
//   B operand = the signal itself.  A tile's N + P columns sit in shared memory as 16-byte
//       core-matrix rows ordered [s/8][column][s%8] (no-swizzle K-major canonical layout), so the
//       "-p" column shift of block p is just +16 bytes on the descriptor start address: X is staged
//       ONCE per tile and read by all P+1 blocks.
//   A operand = Toeplitz block.  With the accumulator rows reversed (r' = 127 - r) the block is
//       Hankel, A[r', s] = g[r' + s + const], so an 8x8 core matrix depends only on r'/8 + s/8.
//       Setting the descriptor's leading AND stride byte offsets to 128 B makes every core matrix
//       of every T_p alias into ONE 8x-expanded tap array H[u][i][e] = g[8u + i + e]
//       ((16 P + 31) * 128 B per split term): no per-block A traffic at all.
//   precision = split 16-bit operands, FP32 accumulation in TMEM.  Two formats:
//       fmt 0 (default) FP16, block-scaled: every slab (tile + halo) is multiplied by a power of two
//           that puts its max |x| in [2^14, 2^15), the taps likewise (once); x' = xh + xm, c' = ch + cm
//           with 11-bit significands, products hh + hm + mh; what is dropped (mm and the two
//           third-order residuals) is <= 3 * 2^-22 per product relative to max|x| * |c| -- an ABSOLUTE
//           bound in exactly the unit of the path's tolerance (1e-5 * sum|h| * max|x|).  The scale is
//           undone (exactly) in the epilogue.
//       fmt 1 BF16 (8-bit significands): 3 / 4 / 6 products (hh, hm, mh, + mm, + hl, lh); kept for A/B.
//
// Warp roles (one persistent CTA per SM, 448 threads):
//   warps 0-3  epilogue: tcgen05.ld their TMEM lane quarter, un-scale, coalesced 128-B row stores to HBM
//   warps 4-11 loader / converter (256 threads)
//   warp  12   MMA     : one elected lane issues tcgen05.mma (cta_group::1, kind::f16, M128 N128 K16)
//   warp  13   TMA     : in-place mode only: one lane prefetches whole fp32 tiles with cp.async.bulk
// Two staging modes:
//   IN-PLACE (K <= 385): three unified 66 KB buffers.  A buffer is filled with the raw fp32 tile by ONE
//       bulk copy (tile t+2 is in flight while tile t+1 is converted and tile t is multiplied), the
//       converter warps pull it into registers, meet at a named barrier (which also carries the
//       max-|x| reduction), and write the split 16-bit slab back INTO THE SAME buffer; the MMA's
//       tcgen05.commit hands the buffer back to the TMA warp.
//   REGISTER (long filters; the Hankel arrays leave room for one or two slabs only): LDG.128 of the
//       NEXT tile into registers while the tensor core still reads the slab, convert + STS.128 as soon
//       as tcgen05.commit frees it.
// Pipelines: buffer raw-full / slab-full / empty, accumulator full / empty (2 x 128 TMEM columns), all
// on mbarriers.
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace scir_b200 {

namespace {

constexpr int TB = 128;                 // block length: MMA M, and the K extent of one p-block
constexpr int TN = 128;                 // blocks (columns) per tile: MMA N (the default; q.tn = 64 trades MMA width for deeper prefetch)
constexpr int kEpiWarps = 4, kLoadWarps = 16;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kMmaWarp = kEpiWarps + kLoadWarps, kTmaWarp = kMmaWarp + 1;
constexpr int kToepThreads = (kEpiWarps + kLoadWarps + 2) * 32;
constexpr int kMaxBuf = 6;
constexpr int kPmaxLimit = 32;
constexpr int RQMAX = ((TN + kPmaxLimit) * 32 + kLoadThreads - 1) / kLoadThreads;  // float4 chunks per loader thread

enum { MODE_REGISTER = 0, MODE_INPLACE = 1 };
enum { FMT_F16_SCALED = 0, FMT_BF16 = 1 };

struct ToepParams {
    FirPass p;
    long long first_col;                // block index of tile 0's first output column
    long long org;                      // block j covers causal indices [128 j - org, 128 j - org + 128): org in [0,3] makes 8-sample chunks 16-B aligned in memory
    long long ip_lo, ip_hi;             // outputs wanted, in the kernel's causal index i' (see map_index)
    long long fast_lo, fast_hi;         // virtual indices i whose sample is x[i + in_off] verbatim
    int fast_ok;                        // 16-byte alignment of 8-sample chunks holds on the fast path
    int tiles_per_row;
    int total_tiles;
    int pmax;                           // last Toeplitz block index P
    int slab_cols;                      // odd, >= TN + pmax
    int nver;                           // split terms kept per operand: 2 (hi, mid) or 3 (+ lo)
    int terms;                          // 3, 4 or 6 products
    int mma_per_tile;                   // MMAs issued per 128 x tn tile (statistic)
    int mode;                           // MODE_REGISTER / MODE_INPLACE
    int nbuf;                           // slab buffers: 3 in place, 1 or 2 in register mode
    int fmt;                            // FMT_F16_SCALED / FMT_BF16
    int tn;                             // columns per tile = MMA N: 128, or 64 (twice the buffers: deeper TMA prefetch)
    int chains;                         // independent accumulation chains per tile (1 or 2): accumulators = 2 stages * chains * 128 columns
    int ts_blocks;                      // Toeplitz blocks T_0 .. T_{ts_blocks-1} are kept in TMEM (A operand from TMEM)
    int tmem_cols;                      // TMEM allocation (power of two): accumulators, then the A blocks
    int k;
    int hank_cores;                     // 16 * pmax + 31
    unsigned buf_bytes;                 // one buffer (multiple of 128)
    float tap_scale, tap_inv;           // power of two applied to the taps (fmt 0) and its inverse
    int* tile_flags;                    // [total_tiles]: 1 if the tile's slab held a NaN / Inf sample (see toeplitz_fixup_kernel)
    const float* taps;                  // [k] by delay index, in device memory (ctx-owned, re-uploaded only when the filter
                                        // changes): keeps the launch parameters at ~300 B -- 32 KB of by-value taps made the
                                        // two launches of a pass cost 30-60 us of host time
    int stream_stores;                  // 1: epilogue stores carry the evict-first hint (st.global.cs): outputs are written once
    unsigned long long hh_mask;         // bit pb set: Toeplitz block pb multiplies hi x hi only (its taps are so small that the two
                                        // mid-term products stay inside the error budget: make_plan, "precision per block")
};

// ---- tcgen05 / TMEM PTX ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Shared-memory matrix descriptors, SWIZZLE_NONE, K-major: 8 rows x 16 B core matrices.  Bits: [0,14)
// addr>>4, [16,30) leading byte offset>>4 (between the two 16-B K chunks of one K=16 MMA), [32,46) stride
// byte offset>>4 (between 8-row groups in M/N), [46,48) version = 1.  Only the low word changes between
// MMAs (start address), so the issue loop does 32-bit adds and the 64-bit value is assembled in PTX.
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);           // SBO = 128 B for both operands
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo)
{
    return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
}
// A operand from TMEM (lane = accumulator row, 16-bit K elements packed two per 32-bit column), B from shared memory
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& lo, const uint4& hi)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(lo.x), "r"(lo.y),
                 "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// kind::f16 instruction descriptor: D = F32 (bit 4), A/B format at bits 7 / 10 (0 = F16, 1 = BF16), both
// K-major, N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc(int fmt, int tn)
{
    const uint32_t ab = (fmt == FMT_BF16) ? 1u : 0u;
    return (1u << 4) | (ab << 7) | (ab << 10) | (static_cast<uint32_t>(tn >> 3) << 17) | (static_cast<uint32_t>(TB >> 4) << 24);
}

// ---- virtual input sequence (same rules as fir_direct.cu: zero / held boundary, odd / even / const ext) ----
__device__ __forceinline__ float tload(const FirPass& p, const float* __restrict__ xr, long long i)
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

// The kernel always runs a CAUSAL filter over u[i']; an anticausal pass is the causal one on the
// time-reversed sequence: u[i'] = v[n_v - 1 - i'], out[n_v - 1 - i'] = out'[i'].
__device__ __forceinline__ long long map_index(const FirPass& p, long long ip)
{
    return (p.dir > 0) ? ip : (p.n_v - 1 - ip);
}

// 4 consecutive samples -> half of a 16-byte core-matrix row per split term (hi | mid | lo), 16-bit each.
// `v` is in causal order already.
template <bool F16, int NVER>
__device__ __forceinline__ void split_store4(const float (&v)[4], float scale, uint32_t dst, uint32_t ver_bytes)
{
    uint32_t hi[2], mid[2], lo[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        if constexpr (F16) {
            const float s0 = v[2 * e] * scale, s1 = v[2 * e + 1] * scale;
            const __half2 h = __floats2half2_rn(s0, s1);
            const float r0 = s0 - __low2float(h), r1 = s1 - __high2float(h);
            const __half2 m = __floats2half2_rn(r0, r1);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            mid[e] = *reinterpret_cast<const uint32_t*>(&m);
            if constexpr (NVER > 2) {
                const __half2 l = __floats2half2_rn(r0 - __low2float(m), r1 - __high2float(m));
                lo[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
        } else {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            const float r0 = v[2 * e] - __low2float(h), r1 = v[2 * e + 1] - __high2float(h);
            const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            mid[e] = *reinterpret_cast<const uint32_t*>(&m);
            if constexpr (NVER > 2) {
                const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - __low2float(m), r1 - __high2float(m));
                lo[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
        }
    }
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + ver_bytes), "r"(mid[0]), "r"(mid[1]) : "memory");
    if constexpr (NVER > 2)
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + 2 * ver_bytes), "r"(lo[0]), "r"(lo[1]) : "memory");
}

__device__ __forceinline__ uint32_t absmax_bits(const float4& a)
{
    const uint32_t m0 = max(__float_as_uint(a.x) & 0x7fffffffu, __float_as_uint(a.y) & 0x7fffffffu);
    const uint32_t m1 = max(__float_as_uint(a.z) & 0x7fffffffu, __float_as_uint(a.w) & 0x7fffffffu);
    return max(m0, m1);
}

// MMA issue for one tile: all P+1 Toeplitz blocks, every K=16 step that meets a non-zero tap, TERMS split
// products each.  Run by the whole MMA warp with warp-uniform values; only the elected lane issues.  With
// CHAINS = 2 consecutive MMAs alternate between two accumulators (summed in the epilogue), so that no MMA
// waits on the accumulator its predecessor is still writing.
template <int TERMS, int CHAINS>
__device__ __forceinline__ void issue_tile(const ToepParams& q, bool leader, uint32_t dcol, uint32_t a_lo0, uint32_t x_lo0,
                                           uint32_t hank16, uint32_t ver16, uint32_t idesc, uint32_t a_tmem0)
{
    constexpr uint32_t NVER = (TERMS == 6) ? 3u : 2u;
    uint32_t idx = 0;
    auto dst = [&]() { return (CHAINS == 2) ? (dcol + (idx & 1u) * TN) : dcol; };
    auto acc = [&]() { return (idx >= static_cast<uint32_t>(CHAINS)) ? 1u : 0u; };
    auto mma = [&](uint32_t a, uint32_t x) {              // both operands from shared memory
        tc_mma(dst(), a, x, idesc, acc());
        ++idx;
    };
    auto mma_ts = [&](uint32_t at, uint32_t x) {          // Toeplitz block from TMEM
        tc_mma_ts(dst(), at, x, idesc, acc());
        ++idx;
    };
    const uint32_t two_cols = 2u * static_cast<uint32_t>(q.slab_cols);
    for (int pb = 0; pb <= q.pmax; ++pb) {
        const int ks0 = max(0, TB * pb - (q.k - 1)) >> 4;                // first K step with a non-zero tap in T_p
        uint32_t a = a_lo0 + 8u * static_cast<uint32_t>(16 * (q.pmax - pb) + 2 * ks0);
        uint32_t x = x_lo0 + static_cast<uint32_t>(ks0) * two_cols + static_cast<uint32_t>(q.pmax - pb);
        const bool full = !((q.hh_mask >> pb) & 1ull);                    // (only ever cleared for TERMS == 3)
        if (pb < q.ts_blocks) {
            uint32_t at = a_tmem0 + static_cast<uint32_t>(pb) * NVER * 64u + static_cast<uint32_t>(ks0) * 8u;
            for (int ks = ks0; ks < TB / 16; ++ks) {
                if (leader) {
                    mma_ts(at, x);                                       // hi * hi
                    if (full) {
                        mma_ts(at, x + ver16);                           // hi * mid
                        mma_ts(at + 64u, x);                             // mid * hi
                    }
                    if constexpr (TERMS >= 4) mma_ts(at + 64u, x + ver16);
                    if constexpr (TERMS == 6) {
                        mma_ts(at, x + 2u * ver16);                      // hi * lo
                        mma_ts(at + 128u, x);                            // lo * hi
                    }
                } else {
                    idx += full ? TERMS : TERMS - 2;
                }
                at += 8u;
                x += two_cols;
            }
        } else {
            for (int ks = ks0; ks < TB / 16; ++ks) {
                if (leader) {
                    mma(a, x);                                           // hi * hi
                    if (full) {
                        mma(a, x + ver16);                               // hi * mid
                        mma(a + hank16, x);                              // mid * hi
                    }
                    if constexpr (TERMS >= 4) mma(a + hank16, x + ver16);
                    if constexpr (TERMS == 6) {
                        mma(a, x + 2u * ver16);                          // hi * lo
                        mma(a + 2u * hank16, x);                         // lo * hi
                    }
                } else {
                    idx += full ? TERMS : TERMS - 2;
                }
                a += 16u;
                x += two_cols;
            }
        }
    }
}

template <bool F16, int NVER>
__global__ void __launch_bounds__(kToepThreads, 1)
fir_toeplitz_kernel(const __grid_constant__ ToepParams q)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // raw_full[3] slab_full[3] buf_empty[3] acc_full[2] acc_empty[2]
    __shared__ __align__(8) unsigned long long bars[3 * kMaxBuf + 4];      // kMaxBuf = 6
    __shared__ uint32_t tmem_base_holder;
    __shared__ uint32_t red_slots[2][kLoadWarps];          // per-warp max|x| bits, double-buffered by tile parity
    __shared__ float inv_scale_ring[8];                    // loader -> epilogue: 1 / slab scale of tile it & 7 (a loader can be
                                                           // at most 4 tiles ahead of the epilogue: see the barrier chain)

    const FirPass& p = q.p;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                      // warp-uniform for the compiler too
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t hank_bytes = static_cast<uint32_t>(q.hank_cores) * 128u;      // per version
    const uint32_t ver_bytes = 16u * static_cast<uint32_t>(q.slab_cols) * 16u;    // slab bytes per version
    const uint32_t buf0_off = hank_bytes * NVER;
    const uint32_t bar0 = smem_u32(&bars[0]);
    enum { RAW_FULL = 0, SLAB_FULL = kMaxBuf, BUF_EMPTY = 2 * kMaxBuf, ACC_FULL = 3 * kMaxBuf, ACC_EMPTY = 3 * kMaxBuf + 2 };
    auto BAR = [&](int which, int idx) { return bar0 + 8u * static_cast<uint32_t>(which + idx); };
    const int tile_cols = q.tn + q.pmax;
    const uint32_t raw_bytes = static_cast<uint32_t>(tile_cols) * TB * 4u;        // one fp32 tile incl. halo columns
    const uint32_t tmem_cols = static_cast<uint32_t>(q.tmem_cols);
    const uint32_t a_col0 = 2u * TN * static_cast<uint32_t>(q.chains);             // A blocks follow the accumulators
    // a tile whose N + P columns are plain, aligned memory can be fetched by one bulk copy
    auto tile_lo = [&](long long ipA) { return (p.dir > 0) ? ipA : (p.n_v - ipA - static_cast<long long>(tile_cols) * TB); };
    auto tile_bulk = [&](long long ipA) {
        const long long lo = tile_lo(ipA);
        return q.fast_ok && lo >= q.fast_lo && lo + static_cast<long long>(tile_cols) * TB <= q.fast_hi;
    };
    auto tile_ipA = [&](int ct) { return (q.first_col + static_cast<long long>(ct) * q.tn - q.pmax) * TB - q.org; };

    // ---- one-time set-up: barriers, TMEM, the 8x-expanded (Hankel) tap arrays ------------------------------
    if (tid == 0) {
        for (int i = 0; i < kMaxBuf; ++i) {
            mbar_init(BAR(RAW_FULL, i), 1);
            mbar_init(BAR(SLAB_FULL, i), kLoadThreads);
            mbar_init(BAR(BUF_EMPTY, i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(ACC_FULL, i), 1);
            mbar_init(BAR(ACC_EMPTY, i), kEpiWarps * 32);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_base_holder), tmem_cols);
    {
        // H[u][i][e] = g[8u + i + e],  g[w] = c[128 P + 127 - w]  (zero outside [0, k))
        const int top = TB * q.pmax + (TB - 1);
        uint16_t* H = reinterpret_cast<uint16_t*>(smem_raw);
        const uint32_t hank_elems = hank_bytes >> 1;
        for (int idx = tid; idx < q.hank_cores * 64; idx += kToepThreads) {
            const int u = idx >> 6, i = (idx >> 3) & 7, e = idx & 7;
            const int ci = top - (8 * u + i + e);
            const float c = (ci >= 0 && ci < q.k) ? __ldg(q.taps + ci) : 0.f;
            uint16_t h16, m16, l16;
            if constexpr (F16) {
                const float cs = c * q.tap_scale;
                const __half h = __float2half_rn(cs);
                const float r1 = cs - __half2float(h);
                const __half m = __float2half_rn(r1);
                const __half l = __float2half_rn(r1 - __half2float(m));
                h16 = __half_as_ushort(h); m16 = __half_as_ushort(m); l16 = __half_as_ushort(l);
            } else {
                const __nv_bfloat16 h = __float2bfloat16_rn(c);
                const float r1 = c - __bfloat162float(h);
                const __nv_bfloat16 m = __float2bfloat16_rn(r1);
                const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
                h16 = __bfloat16_as_ushort(h); m16 = __bfloat16_as_ushort(m); l16 = __bfloat16_as_ushort(l);
            }
            H[idx] = h16;
            H[hank_elems + idx] = m16;
            if constexpr (NVER > 2) H[2 * hank_elems + idx] = l16;
        }
    }
    fence_proxy_async_smem();                              // H is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;
    if (q.ts_blocks > 0) {
        // The first Toeplitz blocks move into TMEM once (same Hankel aliasing, read with LDS.128): their MMAs
        // then fetch only the signal operand from shared memory, which halves the tensor core's load on the
        // shared-memory port -- the resource this kernel is bound by.
        if (warp < kEpiWarps) {
            const int rq = warp * 32 + lane;               // accumulator row r' = TMEM lane
            const int mi = rq >> 3, i = rq & 7;
            for (int pb = 0; pb < q.ts_blocks; ++pb)
                for (int ver = 0; ver < NVER; ++ver)
                    for (int ks = 0; ks < TB / 16; ++ks) {
                        const unsigned char* src = smem_raw + static_cast<size_t>(ver) * hank_bytes +
                                                   static_cast<size_t>(16 * (q.pmax - pb) + 2 * ks + mi) * 128 + i * 16;
                        const uint4 lo = *reinterpret_cast<const uint4*>(src);          // K chunk 0: s = 16 ks + [0, 8)
                        const uint4 hi = *reinterpret_cast<const uint4*>(src + 128);    // K chunk 1: s = 16 ks + [8, 16)
                        tmem_st8(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + a_col0 +
                                     static_cast<uint32_t>((pb * NVER + ver) * 64 + ks * 8),
                                 lo, hi);
                    }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    const int ntiles = q.total_tiles;
    const int nbuf = q.nbuf;
    if (warp == kTmaWarp) {
        // ===== TMA PRODUCER (in-place mode): whole fp32 tiles by cp.async.bulk, as soon as a buffer is free ===
        if (q.mode == MODE_INPLACE && lane == 0) {
            int buf = 0;
            uint32_t par = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(BAR(BUF_EMPTY, buf), par ^ 1u);
                const int row = t / q.tiles_per_row;
                const int ct = t - row * q.tiles_per_row;
                const long long ipA = tile_ipA(ct);
                if (tile_bulk(ipA)) {
                    mbar_arrive_expect_tx(BAR(RAW_FULL, buf), raw_bytes);
                    bulk_copy_g2s(smem0 + buf0_off + static_cast<uint32_t>(buf) * q.buf_bytes,
                                  p.x + static_cast<long long>(row) * p.ld_x + tile_lo(ipA) + p.in_off, raw_bytes, BAR(RAW_FULL, buf));
                } else {
                    mbar_arrive(BAR(RAW_FULL, buf));       // edge tile: the converters synthesise it from global memory
                }
                if (++buf == nbuf) { buf = 0; par ^= 1u; }
            }
        }
    } else if (warp >= kEpiWarps && warp < kEpiWarps + kLoadWarps) {
        // ===== LOADER / CONVERTER: stage the tile's N + P columns as split 16-bit core-matrix rows ============
        // A thread owns the float4 chunks tl, tl + 512, ... of the tile (chunk u = causal samples [4u, 4u+4) of
        // the slab = half `u & 1` of core-matrix row (column u >> 5, K chunk (u >> 1) & 15)): lane-contiguous
        // LDS.128 / LDG.128 in, conflict-free STS.64 out.
        const int tl = tid - kEpiWarps * 32;
        const int lw = tl >> 5;
        const int nq = tile_cols * 32;
        const bool inplace = (q.mode == MODE_INPLACE);
        float4 raw[RQMAX];
        uint32_t fastmask = 0;                             // chunks held in ascending memory order (anticausal: to be reversed)
        uint32_t mx = 0;
        // gather from global memory: aligned interior chunks by LDG.128, edge chunks rule by rule
        auto fetch = [&](int t) {
            const int row = t / q.tiles_per_row;
            const int ct = t - row * q.tiles_per_row;
            const long long ipA = tile_ipA(ct);
            const float* __restrict__ xr = p.x + static_cast<long long>(row) * p.ld_x;
            fastmask = 0;
#pragma unroll
            for (int r = 0; r < RQMAX; ++r) {
                const int u = tl + r * kLoadThreads;
                if (u < nq) {
                    const long long ip0 = ipA + 4 * u;                         // causal index of the chunk's first sample
                    // virtual index of the chunk's lowest address: ip0 (causal) or n_v-1-ip0-3 (anticausal)
                    const long long lo = (p.dir > 0) ? ip0 : (p.n_v - 4 - ip0);
                    if (ip0 >= q.ip_hi) {                                      // newer than every wanted output: never read
                        raw[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else if (q.fast_ok && lo >= q.fast_lo && lo + 4 <= q.fast_hi) {
                        raw[r] = *reinterpret_cast<const float4*>(xr + lo + p.in_off);
                        fastmask |= 1u << r;
                    } else {                                                   // edge chunk: causal order, rule by rule
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = tload(p, xr, (p.dir > 0) ? (ip0 + e) : (p.n_v - 1 - ip0 - e));
                        raw[r] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
        };
        // gather from a raw fp32 tile in shared memory (ascending memory order)
        auto gather_smem = [&](const float* rawt) {
            const int len = tile_cols * TB;
            const bool rev = p.dir < 0;
            fastmask = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < RQMAX; ++r) {
                const int u = tl + r * kLoadThreads;
                if (u < nq) raw[r] = *reinterpret_cast<const float4*>(rev ? (rawt + (len - 4 - 4 * u)) : (rawt + 4 * u));
            }
        };
        auto local_max = [&]() {
            mx = 0;
#pragma unroll
            for (int r = 0; r < RQMAX; ++r)
                if (tl + r * kLoadThreads < nq) mx = max(mx, absmax_bits(raw[r]));
        };
        // block-wide max -> power-of-two scale (fmt 0); the named barrier doubles as the "everyone has
        // read the raw tile" point of the in-place conversion
        auto slab_scale = [&](int it, int t) -> float {
            // One reduction per slab, both formats: a non-finite sample cannot be represented by the split (and
            // 0 * NaN would spread it over its whole 128-sample block), so such tiles are flagged and redone
            // exactly by toeplitz_fixup_kernel.  In FP16 mode the max also sets the block scale.
            const uint32_t wm = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0) red_slots[it & 1][lw] = wm;
            asm volatile("bar.sync 1, %0;" ::"n"(kLoadThreads) : "memory");
            uint32_t m = *(volatile uint32_t*)&red_slots[it & 1][lane & (kLoadWarps - 1)];
            m = __reduce_max_sync(0xffffffffu, m);
            if (tl == 0) q.tile_flags[t] = (m >= 0x7f800000u) ? 1 : 0;
            if constexpr (!F16) {
                return 1.f;
            } else {
                // scale = 2^(141 - E): max|x| * scale in [2^14, 2^15); exponent field kept in [14, 253] so that
                // both the scale and its inverse are normal numbers
                const int e = static_cast<int>(m >> 23);
                const int sexp = min(268 - e, 253);
                if (tl == 0) inv_scale_ring[it & 7] = __uint_as_float(static_cast<uint32_t>(254 - sexp) << 23);
                return __uint_as_float(static_cast<uint32_t>(sexp) << 23);
            }
        };
        auto store_slab = [&](uint32_t sbase, float scale) {
            const uint32_t row_bytes = static_cast<uint32_t>(q.slab_cols) * 16u;
            if (p.dir > 0) {
#pragma unroll
                for (int r = 0; r < RQMAX; ++r) {
                    const int u = tl + r * kLoadThreads;
                    if (u < nq) {
                        const float v[4] = {raw[r].x, raw[r].y, raw[r].z, raw[r].w};
                        const uint32_t dst = sbase + static_cast<uint32_t>((u >> 1) & 15) * row_bytes + static_cast<uint32_t>(u >> 5) * 16u + (u & 1) * 8u;
                        split_store4<F16, NVER>(v, scale, dst, ver_bytes);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < RQMAX; ++r) {
                    const int u = tl + r * kLoadThreads;
                    if (u < nq) {
                        const bool rv = (fastmask >> r) & 1u;                  // ascending-memory chunks are time-reversed
                        const float v[4] = {rv ? raw[r].w : raw[r].x, rv ? raw[r].z : raw[r].y, rv ? raw[r].y : raw[r].z,
                                            rv ? raw[r].x : raw[r].w};
                        const uint32_t dst = sbase + static_cast<uint32_t>((u >> 1) & 15) * row_bytes + static_cast<uint32_t>(u >> 5) * 16u + (u & 1) * 8u;
                        split_store4<F16, NVER>(v, scale, dst, ver_bytes);
                    }
                }
            }
        };

        int it = 0, buf = 0;
        uint32_t par = 0;
        if (inplace) {
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int row = t / q.tiles_per_row;
                const int ct = t - row * q.tiles_per_row;
                const uint32_t boff = buf0_off + static_cast<uint32_t>(buf) * q.buf_bytes;
                mbar_wait(BAR(RAW_FULL, buf), par);        // bulk copy landed (or: buffer is free, edge tile)
                if (tile_bulk(tile_ipA(ct))) gather_smem(reinterpret_cast<const float*>(smem_raw + boff));
                else fetch(t);
                local_max();
                const float scale = slab_scale(it, t);     // barrier: every converter holds its part of the tile
                store_slab(smem0 + boff, scale);
                fence_proxy_async_smem();                  // my generic-proxy writes -> visible to the MMA's async reads
                mbar_arrive(BAR(SLAB_FULL, buf));
                if (++buf == nbuf) { buf = 0; par ^= 1u; }
            }
        } else {
            // the thread's share of the NEXT tile is fetched into registers right after the current one is
            // stored, so the HBM latency is spent while the tensor core works and only the convert +
            // STS phase waits on buf_empty
            if (static_cast<int>(blockIdx.x) < ntiles) fetch(blockIdx.x);
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                local_max();
                const float scale = slab_scale(it, t);
                mbar_wait(BAR(BUF_EMPTY, buf), par ^ 1u);
                store_slab(smem0 + buf0_off + static_cast<uint32_t>(buf) * q.buf_bytes, scale);
                fence_proxy_async_smem();
                mbar_arrive(BAR(SLAB_FULL, buf));
                if (t + static_cast<int>(gridDim.x) < ntiles) fetch(t + gridDim.x);
                if (++buf == nbuf) { buf = 0; par ^= 1u; }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA ISSUER ===================================================================================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc(q.fmt, q.tn);
        const uint32_t a_lo0 = desc_lo(smem0, 128u);
        const uint32_t hank16 = hank_bytes >> 4, ver16 = ver_bytes >> 4;
        const uint32_t lbo_x = static_cast<uint32_t>(q.slab_cols) * 16u;
        const uint32_t acc_cols = TN * static_cast<uint32_t>(q.chains);
        int it = 0, buf = 0;
        uint32_t par = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t apar = (it >> 1) & 1;
            mbar_wait(BAR(ACC_EMPTY, acc), apar ^ 1u);
            mbar_wait(BAR(SLAB_FULL, buf), par);
            tc_fence_after();
            const uint32_t x_lo0 = desc_lo(smem0 + buf0_off + static_cast<uint32_t>(buf) * q.buf_bytes, lbo_x);
            const uint32_t dcol = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
            constexpr int T0 = (NVER > 2) ? 6 : 3;
            const uint32_t a_tmem0 = tmem_base + a_col0;
            if (q.chains == 2) {
                if (NVER == 2 && q.terms == 4) issue_tile<4, 2>(q, leader, dcol, a_lo0, x_lo0, hank16, ver16, idesc, a_tmem0);
                else issue_tile<T0, 2>(q, leader, dcol, a_lo0, x_lo0, hank16, ver16, idesc, a_tmem0);
            } else {
                if (NVER == 2 && q.terms == 4) issue_tile<4, 1>(q, leader, dcol, a_lo0, x_lo0, hank16, ver16, idesc, a_tmem0);
                else issue_tile<T0, 1>(q, leader, dcol, a_lo0, x_lo0, hank16, ver16, idesc, a_tmem0);
            }
            if (leader) {
                tc_commit(BAR(BUF_EMPTY, buf));            // buffer may be refilled once these MMAs retire
                tc_commit(BAR(ACC_FULL, acc));             // accumulator(s) complete
            }
            __syncwarp();
            if (++buf == nbuf) { buf = 0; par ^= 1u; }
        }
    } else {
        // ===== EPILOGUE: TMEM -> registers -> coalesced row stores ==========================================
        int it = 0;
        const int rr = TB - 1 - (warp * 32 + lane);        // accumulator rows are reversed (Hankel trick)
        const uint32_t acc_cols = TN * static_cast<uint32_t>(q.chains);
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t apar = (it >> 1) & 1;
            const int row = t / q.tiles_per_row;
            const int ct = t - row * q.tiles_per_row;
            const long long j0 = q.first_col + static_cast<long long>(ct) * q.tn;
            float* __restrict__ yr = p.y + static_cast<long long>(row) * p.ld_y + p.out_off;
            mbar_wait(BAR(ACC_FULL, acc), apar);
            tc_fence_after();
            // un-scale: exact powers of two (1 in BF16 mode); written by the loaders long before ACC_FULL
            const float inv_x = F16 ? *(volatile float*)&inv_scale_ring[it & 7] : 1.f;
            const float inv_c = q.tap_inv;
            const long long ip_first = j0 * TB - q.org, ip_last = (j0 + q.tn) * TB - q.org;      // this tile's outputs [first, last)
            const bool interior = ip_first >= q.ip_lo && ip_last <= q.ip_hi;
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(acc) * acc_cols;
#pragma unroll 1
            for (int c4 = 0; c4 < q.tn / 32; ++c4) {
                uint32_t v[32];
                tmem_ld32(tbase + static_cast<uint32_t>(c4 * 32), v);
                if (q.chains == 2) {                       // second accumulation chain
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t w[16];
                        tmem_ld16(tbase + static_cast<uint32_t>(TN + c4 * 32 + hh * 16), w);
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            v[hh * 16 + c] = __float_as_uint(__uint_as_float(v[hh * 16 + c]) + __uint_as_float(w[c]));
                    }
                }
                const long long ipb = (j0 + c4 * 32) * TB + rr - q.org;
                if (interior) {                            // one base pointer, immediate offsets, no guards
                    float* yb = yr + map_index(p, ipb);
                    if (q.stream_stores) {
                        if (p.dir > 0) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) __stcs(yb + c * TB, __uint_as_float(v[c]) * inv_x * inv_c);
                        } else {
#pragma unroll
                            for (int c = 0; c < 32; ++c) __stcs(yb - c * TB, __uint_as_float(v[c]) * inv_x * inv_c);
                        }
                    } else if (p.dir > 0) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) yb[c * TB] = __uint_as_float(v[c]) * inv_x * inv_c;
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) yb[-c * TB] = __uint_as_float(v[c]) * inv_x * inv_c;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const long long ip = ipb + static_cast<long long>(c) * TB;
                        if (ip >= q.ip_lo && ip < q.ip_hi) yr[map_index(p, ip)] = __uint_as_float(v[c]) * inv_x * inv_c;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(BAR(ACC_EMPTY, acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Tiles whose slab held a NaN / Inf sample are recomputed here the way the FP32 direct kernels (and the
// reference loop, crates/scir-gpu/src/lib.rs:1138-1150) do it: one FMA chain per output, so the non-finite value
// reaches exactly the outputs whose window contains it.  Launched after every Toeplitz pass; with finite
// data it reads total_tiles flags from L2 and exits (~3 us).
__global__ void __launch_bounds__(256) toeplitz_fixup_kernel(const __grid_constant__ ToepParams q)
{
    const FirPass& p = q.p;
    // 256 flags per CTA trip, one per thread (coalesced); the common case is one load and one barrier
    for (int base = blockIdx.x * 256; base < q.total_tiles; base += gridDim.x * 256) {
        const int mine = base + threadIdx.x;
        const int flag = (mine < q.total_tiles) ? q.tile_flags[mine] : 0;
        if (!__syncthreads_or(flag)) continue;
        __shared__ int hit[256];
        hit[threadIdx.x] = flag;
        __syncthreads();
        for (int j = 0; j < 256; ++j) {
            if (!hit[j]) continue;
            const int t = base + j;
            const int row = t / q.tiles_per_row;
            const int ct = t - row * q.tiles_per_row;
            const long long ip0 = (q.first_col + static_cast<long long>(ct) * q.tn) * TB - q.org;
            const float* __restrict__ xr = p.x + static_cast<long long>(row) * p.ld_x;
            float* __restrict__ yr = p.y + static_cast<long long>(row) * p.ld_y + p.out_off;
            for (int o = threadIdx.x; o < TB * q.tn; o += blockDim.x) {
                const long long ip = ip0 + o;
                if (ip < q.ip_lo || ip >= q.ip_hi) continue;
                const long long i = map_index(p, ip);                  // virtual index of this output
                float acc = 0.f;
                for (int d = 0; d < q.k; ++d) acc = fmaf(__ldg(q.taps + d), tload(p, xr, p.dir > 0 ? i - d : i + d), acc);
                yr[i] = acc;
            }
        }
        __syncthreads();
    }
}

struct ToepPlan {
    ToepParams q;
    size_t smem_bytes;
};

bool make_plan(const scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k, ToepPlan* out)
{
    if (k < 1 || k > SCIR_B200_MAX_TAPS) return false;
    ToepParams q{};
    q.p = pass;
    q.k = static_cast<int>(k);
    q.pmax = static_cast<int>((k - 1 + (TB - 1)) / TB);
    if (q.pmax > kPmaxLimit) return false;                 // loader register budget (RMAX) and shared memory
    q.hank_cores = 16 * q.pmax + 31;
    // tile width: 64 columns give twice as many (half-size) buffers for the in-place pipeline, i.e. a deeper TMA
    // prefetch, at half the MMA width: only for filters short enough to be HBM-bound (toeplitz_tn: 0 auto, 64, 128)
    q.tn = (ctx->opt.toeplitz_tn == 64) ? 64 : (ctx->opt.toeplitz_tn == 128) ? TN : ((q.pmax <= 1) ? ctx->opt.toeplitz_tn_short : TN);
    if (q.tn != 64) q.tn = TN;
    q.slab_cols = (q.tn + q.pmax) | 1;
    q.fmt = (ctx->opt.toeplitz_split == 1) ? FMT_BF16 : FMT_F16_SCALED;
    int64_t terms = ctx->opt.toeplitz_terms;
    if (terms != 3 && terms != 4 && terms != 6) terms = 3;
    q.terms = static_cast<int>(terms);
    q.nver = (terms == 6) ? 3 : 2;
    q.chains = (ctx->opt.toeplitz_chains == 2) ? 2 : 1;
    q.stream_stores = (ctx->opt.toeplitz_stcs != 0) ? 1 : 0;
    {   // TMEM: 2 accumulator stages x chains x 128 columns, then as many Toeplitz blocks (64 columns per split term)
        // as fit in the 512 columns; blocks beyond that stay shared-memory operands
        const int acc_cols = 2 * TN * q.chains;
        const int room = (512 - acc_cols) / (64 * q.nver);
        q.ts_blocks = (ctx->opt.toeplitz_ts == 0) ? 0 : std::min(q.pmax + 1, room);
        const int need = acc_cols + q.ts_blocks * 64 * q.nver;
        q.tmem_cols = 32;
        while (q.tmem_cols < need) q.tmem_cols *= 2;
    }
    const size_t hank = static_cast<size_t>(q.hank_cores) * 128 * q.nver;
    const size_t slab = static_cast<size_t>(16) * q.slab_cols * 16 * q.nver;
    const size_t raw = static_cast<size_t>(q.tn + q.pmax) * TB * 4;
    const size_t budget = static_cast<size_t>(ctx->max_smem_optin) - 1024;        // static smem + slack
    auto up128 = [](size_t v) { return (v + 127) / 128 * 128; };
    if (ctx->opt.toeplitz_loader != 1 && hank + 3 * up128(std::max(slab, raw)) <= budget) {
        q.mode = MODE_INPLACE;                             // short filters: TMA-fed buffers converted in place
        q.buf_bytes = static_cast<unsigned>(up128(std::max(slab, raw)));
        q.nbuf = static_cast<int>(std::min<size_t>(kMaxBuf, (budget - hank) / q.buf_bytes));
    } else {
        q.mode = MODE_REGISTER;                            // long filters: register prefetch
        q.buf_bytes = static_cast<unsigned>(up128(slab));
        if (hank + q.buf_bytes > budget) return false;
        q.nbuf = (hank + 2 * static_cast<size_t>(q.buf_bytes) <= budget) ? 2 : 1;
    }
    // FP16 block scaling of the taps: max|c| * tap_scale in [2^14, 2^15)
    q.tap_scale = 1.f;
    q.tap_inv = 1.f;
    if (q.fmt == FMT_F16_SCALED && c != nullptr) {
        float cmax = 0.f;
        for (int64_t i = 0; i < k; ++i) cmax = std::max(cmax, std::fabs(c[i]));
        if (cmax > 0.f && std::isfinite(cmax)) {
            int e = 0;
            std::frexp(cmax, &e);                          // cmax = m * 2^e, m in [0.5, 1)
            const int se = std::min(std::max(15 - e, -120), 120);
            q.tap_scale = std::ldexp(1.f, se);
            q.tap_inv = std::ldexp(1.f, -se);
        }
    }
    // ---- precision per Toeplitz block ------------------------------------------------------------------------------
    // Block pb holds the taps c[128 pb - 127 .. 128 pb + 127].  Dropping its two mid-term products (c_h x_m + c_m x_h)
    // costs at most 2 * 2^-11 * S_pb * max|x| per output, S_pb = sum of |c| over those taps (|x_m| <= 2^-11 |x| and
    // |c_m| <= 2^-11 |c|: FP16 keeps 11 significant bits).  Windowed-sinc designs put almost all of sum|c| in one or two
    // blocks: the others are multiplied hi x hi only, as long as the dropped terms of ALL such blocks together stay
    // within `toeplitz_adaptive_budget` thousandths of the path's tolerance 1e-5 * sum|c| * max|x| (default 0.5; the
    // always-dropped mid x mid products, the operands' third terms and FP32 accumulation use another ~0.15-0.2 in the worst
    // case).  Config 5's fused 509-tap pass: the two outer of its 5 blocks hold 0.45 % of sum|c|: 120 -> 88 MMAs per tile.  Random or flat taps: nothing is dropped.  0 disables.
    q.hh_mask = 0ull;
    q.mma_per_tile = 0;
    {
        std::vector<std::pair<double, int>> blocks;
        double total = 0.0;
        if (c != nullptr)
            for (int64_t i = 0; i < k; ++i) total += std::fabs(static_cast<double>(c[i]));
        if (c != nullptr && q.fmt == FMT_F16_SCALED && q.terms == 3 && ctx->opt.toeplitz_adaptive_budget > 0 && total > 0.0 &&
            std::isfinite(total) && q.pmax < 64) {
            for (int pb = 0; pb <= q.pmax; ++pb) {
                const int64_t lo = std::max<int64_t>(0, static_cast<int64_t>(TB) * pb - (TB - 1));
                const int64_t hi = std::min<int64_t>(k - 1, static_cast<int64_t>(TB) * pb + (TB - 1));
                double sb = 0.0;
                for (int64_t i = lo; i <= hi; ++i) sb += std::fabs(static_cast<double>(c[i]));
                blocks.emplace_back(sb, pb);
            }
            std::sort(blocks.begin(), blocks.end());
            const double budget = static_cast<double>(ctx->opt.toeplitz_adaptive_budget) * 1e-3 * 1e-5 * total / (2.0 * 4.8828125e-4);
            double used = 0.0;
            for (const auto& b : blocks) {
                if (used + b.first > budget) break;
                used += b.first;
                q.hh_mask |= 1ull << b.second;
            }
        }
        for (int pb = 0; pb <= q.pmax; ++pb) {
            const int ks0 = std::max(0, TB * pb - (q.k - 1)) >> 4;
            q.mma_per_tile += (TB / 16 - ks0) * (((q.hh_mask >> pb) & 1ull) ? q.terms - 2 : q.terms);
        }
    }
    // outputs wanted, in causal index space i'
    q.ip_lo = (pass.dir > 0) ? pass.out_begin : (pass.n_v - pass.out_end);
    q.ip_hi = (pass.dir > 0) ? pass.out_end : (pass.n_v - pass.out_begin);
    // origin shift: chunk addresses are (ip0 + in_off) causal, (n_v - 8 - ip0 + in_off) anticausal, ip0 = 8m - org
    {
        const long long al0 = (pass.dir > 0) ? pass.in_off : -(pass.n_v + pass.in_off);
        q.org = ((al0 % 4) + 4) % 4;
    }
    const long long tile_len = static_cast<long long>(TB) * q.tn;
    const long long first_tile = (q.ip_lo + q.org) / tile_len;          // ip + org >= 0 is the block-grid coordinate
    q.first_col = first_tile * q.tn;
    const long long tiles = (q.ip_hi + q.org + tile_len - 1) / tile_len - first_tile;
    if (tiles <= 0 || tiles * pass.batch > 0x7fffffffLL) return false;
    q.tiles_per_row = static_cast<int>(tiles);
    q.total_tiles = static_cast<int>(tiles * pass.batch);
    q.fast_lo = 0;
    q.fast_hi = pass.n_v;
    if (pass.ext_mode != EXT_NONE) {
        q.fast_lo = std::max<long long>(q.fast_lo, -pass.in_off);
        q.fast_hi = std::min<long long>(q.fast_hi, pass.n_x - pass.in_off);
    }
    q.fast_ok = aligned16(pass.x) && (pass.ld_x % 4 == 0);
    out->q = q;
    out->smem_bytes = hank + static_cast<size_t>(q.nbuf) * q.buf_bytes;
    return true;
}

}  // namespace

bool toeplitz_supported(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k, int64_t* tiles, bool* aligned)
{
    ToepPlan plan;
    if (!(pass.batch > 0 && pass.out_end > pass.out_begin && make_plan(ctx, pass, nullptr, k, &plan))) return false;
    if (tiles) *tiles = static_cast<int64_t>(plan.q.total_tiles) * plan.q.tn / TN;     // in units of 128 x 128 outputs
    if (aligned) *aligned = plan.q.fast_ok != 0;
    return true;
}

int launch_fir_toeplitz(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k)
{
    if (pass.batch <= 0 || pass.out_end <= pass.out_begin) return SCIR_B200_OK;
    ToepPlan plan;
    if (!make_plan(ctx, pass, c, k, &plan))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "tcgen05 Toeplitz path cannot serve k=%lld", (long long)k);
    SCIR_TRY(ctx_bind(ctx));
    // taps live in a ctx-owned device buffer; upload only when the filter changed since the last launch
    SCIR_TRY(ctx_scratch(ctx, ctx->toep_taps, static_cast<size_t>(SCIR_B200_MAX_TAPS) * sizeof(float)));
    if (ctx->toep_taps_host.size() != static_cast<size_t>(k) ||
        std::memcmp(ctx->toep_taps_host.data(), c, static_cast<size_t>(k) * sizeof(float)) != 0) {
        ctx->toep_taps_host.assign(c, c + k);
        // pageable source: the runtime stages it before returning, so `c` may be reused by the caller at once
        SCIR_CUDA(cudaMemcpyAsync(ctx->toep_taps.ptr, ctx->toep_taps_host.data(), static_cast<size_t>(k) * sizeof(float),
                                  cudaMemcpyHostToDevice, ctx->stream),
                  "cudaMemcpyAsync(toeplitz taps)");
    }
    plan.q.taps = static_cast<const float*>(ctx->toep_taps.ptr);
    using Kern = void (*)(const ToepParams);
    const bool f16 = (plan.q.fmt == FMT_F16_SCALED);
    const int flavour = (f16 ? 0 : 2) + (plan.q.nver > 2 ? 1 : 0);
    const Kern kerns[4] = {fir_toeplitz_kernel<true, 2>, fir_toeplitz_kernel<true, 3>, fir_toeplitz_kernel<false, 2>,
                           fir_toeplitz_kernel<false, 3>};
    const Kern kern = kerns[flavour];
    static thread_local size_t configured[16][4] = {};
    const int d = ctx->device & 15;
    if (configured[d][flavour] < plan.smem_bytes) {
        SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.smem_bytes)),
                  "cudaFuncSetAttribute(fir_toeplitz_kernel)");
        configured[d][flavour] = plan.smem_bytes;
    }
    SCIR_TRY(ctx_scratch(ctx, ctx->toep_flags, static_cast<size_t>(plan.q.total_tiles) * sizeof(int)));
    plan.q.tile_flags = static_cast<int*>(ctx->toep_flags.ptr);
    const int grid = std::min(plan.q.total_tiles, ctx->sm_count);
    kern<<<grid, kToepThreads, plan.smem_bytes, ctx->stream>>>(plan.q);
    SCIR_CUDA(cudaGetLastError(), "fir_toeplitz_kernel launch");
    toeplitz_fixup_kernel<<<std::min((plan.q.total_tiles + 255) / 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(plan.q);
    SCIR_CUDA(cudaGetLastError(), "fir_toeplitz_kernel launch");
    ctx->launches += 2;                                    // the contraction and its (normally idle) fix-up
    ctx->fixup_launches++;
    ctx->toeplitz_launches++;
    ctx->toeplitz_last_mma_per_tile = plan.q.mma_per_tile;
    ctx->toeplitz_last_hh_blocks = __builtin_popcountll(plan.q.hh_mask);
    return SCIR_B200_OK;
}

}  // namespace scir_b200
