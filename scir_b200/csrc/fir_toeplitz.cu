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
//   precision = split-BF16: x = xh + xm (+ xl), c = ch + cm (+ cl), products hh, hm, mh (3 terms),
//       + mm (4 terms), + hl, lh (6 terms); FP32 accumulation in TMEM.  Worst case per product
//       3 * 2^-18 (3 terms) / 2 * 2^-18 (4 terms) relative, against the path's 1e-5 tolerance.
//
// Warp roles (one persistent CTA per SM, 448 threads):
//   warps 0-3  epilogue: tcgen05.ld their TMEM lane quarter, coalesced 128-B row stores to HBM
//   warps 4-11 loader  : LDG.128 of the NEXT tile into registers while the tensor core still reads the
//                        slab, then BF16 split -> STS.128 (fence.proxy.async) as soon as the slab is free
//   warp  12   MMA     : one elected lane issues tcgen05.mma (cta_group::1, kind::f16, M128 N128 K16)
//   warp  13   TMA     : short filters only (smem to spare): one lane prefetches whole fp32 tiles two
//                        ahead with cp.async.bulk into a raw ring; warps 4-11 then convert smem -> smem
// Pipelines: slab full/empty (loader <-> MMA, tcgen05.commit releases), accumulator full/empty
// (MMA <-> epilogue, 2 x 128 TMEM columns), all on mbarriers.
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cuda_bf16.h>

namespace scir_b200 {

namespace {

constexpr int TB = 128;                 // block length: MMA M, and the K extent of one p-block
constexpr int TN = 128;                 // blocks (columns) per tile: MMA N
constexpr int kEpiWarps = 4, kLoadWarps = 8;
constexpr int kMmaWarp = kEpiWarps + kLoadWarps, kTmaWarp = kMmaWarp + 1;
constexpr int kToepThreads = (kEpiWarps + kLoadWarps + 2) * 32;
constexpr int kTmemCols = 2 * TN;       // two accumulator stages

struct ToepTaps {
    float c[SCIR_B200_MAX_TAPS];        // by delay index, zero padded
};

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
    int stages;                         // slab stages, 1 or 2
    int raw_stages;                     // 2: fp32 tiles are prefetched by TMA bulk copies into a raw ring; 0: register prefetch
    int k;
    int hank_cores;                     // 16 * pmax + 31
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
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

// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major: 8 rows x 16 B core matrices; `sbo` = byte
// stride between 8-row groups (M/N direction), `lbo` = byte stride between the two 16-B K chunks of
// one K=16 MMA.  Bits: [0,14) addr>>4, [16,30) lbo>>4, [32,46) sbo>>4, [46,48) version = 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = TN.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(TN >> 3) << 17) |
                            (static_cast<uint32_t>(TB >> 4) << 24);

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

__device__ __forceinline__ void split_store(const float (&v)[8], int nver, uint32_t dst, uint32_t ver_bytes)
{
    uint32_t hi[4], mid[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        const float r0 = v[2 * e] - __low2float(h), r1 = v[2 * e + 1] - __high2float(h);
        const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
        const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - __low2float(m), r1 - __high2float(m));
        hi[e] = *reinterpret_cast<const uint32_t*>(&h);
        mid[e] = *reinterpret_cast<const uint32_t*>(&m);
        lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ver_bytes), "r"(mid[0]), "r"(mid[1]), "r"(mid[2]), "r"(mid[3]) : "memory");
    if (nver > 2)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2 * ver_bytes), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
}

__global__ void __launch_bounds__(kToepThreads, 1)
fir_toeplitz_kernel(const __grid_constant__ ToepParams q, const __grid_constant__ ToepTaps taps)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[12];   // slab_full[2] slab_empty[2] acc_full[2] acc_empty[2] raw_full[2] raw_empty[2]
    __shared__ uint32_t tmem_base_holder;

    const FirPass& p = q.p;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t hank_bytes = static_cast<uint32_t>(q.hank_cores) * 128u;      // per version
    const uint32_t ver_bytes = 16u * static_cast<uint32_t>(q.slab_cols) * 16u;    // slab bytes per version
    const uint32_t stage_bytes = ver_bytes * static_cast<uint32_t>(q.nver);
    const uint32_t slab0 = smem0 + hank_bytes * static_cast<uint32_t>(q.nver);
    const uint32_t bar0 = smem_u32(&bars[0]);
    auto BAR = [&](int which, int idx) { return bar0 + 8u * static_cast<uint32_t>(which * 2 + idx); };
    enum { SLAB_FULL = 0, SLAB_EMPTY = 1, ACC_FULL = 2, ACC_EMPTY = 3, RAW_FULL = 4, RAW_EMPTY = 5 };
    const uint32_t raw_bytes = static_cast<uint32_t>(TN + q.pmax) * TB * 4u;     // one fp32 tile incl. halo columns
    const uint32_t raw0_off = (hank_bytes + static_cast<uint32_t>(q.stages) * ver_bytes) * static_cast<uint32_t>(q.nver);
    // a tile whose N + P columns are plain, aligned memory can be fetched by one bulk copy
    auto tile_lo = [&](long long ipA) { return (p.dir > 0) ? ipA : (p.n_v - ipA - static_cast<long long>(TN + q.pmax) * TB); };
    auto tile_bulk = [&](long long ipA) {
        const long long lo = tile_lo(ipA);
        return q.fast_ok && lo >= q.fast_lo && lo + static_cast<long long>(TN + q.pmax) * TB <= q.fast_hi;
    };

    // ---- one-time set-up: barriers, TMEM, the 8x-expanded (Hankel) tap arrays ------------------------------
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(SLAB_FULL, i), kLoadWarps * 32);
            mbar_init(BAR(SLAB_EMPTY, i), 1);
            mbar_init(BAR(ACC_FULL, i), 1);
            mbar_init(BAR(ACC_EMPTY, i), kEpiWarps * 32);
            mbar_init(BAR(RAW_FULL, i), 1);
            mbar_init(BAR(RAW_EMPTY, i), kLoadWarps * 32);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_base_holder), kTmemCols);
    {
        // H[u][i][e] = g[8u + i + e],  g[w] = c[128 P + 127 - w]  (zero outside [0, k))
        const int top = TB * q.pmax + (TB - 1);
        for (int idx = tid; idx < q.hank_cores * 64; idx += kToepThreads) {
            const int u = idx >> 6, i = (idx >> 3) & 7, e = idx & 7;
            const int ci = top - (8 * u + i + e);
            const float c = (ci >= 0 && ci < q.k) ? taps.c[ci] : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(c);
            const float r1 = c - __bfloat162float(h);
            const __nv_bfloat16 m = __float2bfloat16_rn(r1);
            __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(smem_raw);
            H[idx] = h;
            H[(hank_bytes >> 1) + idx] = m;
            if (q.nver > 2) H[2 * (hank_bytes >> 1) + idx] = __float2bfloat16_rn(r1 - __bfloat162float(m));
        }
    }
    fence_proxy_async_smem();                              // H is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;

    const int ntiles = q.total_tiles;
    if (warp == kTmaWarp) {
        // ===== TMA PRODUCER (short filters): whole fp32 tiles, two ahead, by cp.async.bulk ==================
        if (q.raw_stages && lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int rs = it & 1;
                mbar_wait(BAR(RAW_EMPTY, rs), ((it >> 1) & 1) ^ 1u);
                const int row = t / q.tiles_per_row;
                const int ct = t - row * q.tiles_per_row;
                const long long ipA = (q.first_col + static_cast<long long>(ct) * TN - q.pmax) * TB - q.org;
                if (tile_bulk(ipA)) {
                    mbar_arrive_expect_tx(BAR(RAW_FULL, rs), raw_bytes);
                    bulk_copy_g2s(smem0 + raw0_off + static_cast<uint32_t>(rs) * raw_bytes,
                                  p.x + static_cast<long long>(row) * p.ld_x + tile_lo(ipA) + p.in_off, raw_bytes, BAR(RAW_FULL, rs));
                } else {
                    mbar_arrive(BAR(RAW_FULL, rs));        // edge tile: the converters synthesise it
                }
            }
        }
    } else if (warp >= kEpiWarps && warp < kEpiWarps + kLoadWarps && q.raw_stages) {
        // ===== CONVERTER (short filters): raw fp32 tile in smem -> split-BF16 core-matrix rows ===============
        const int tl = tid - kEpiWarps * 32;
        const int len = (TN + q.pmax) * TB;
        const int nitems = (TN + q.pmax) * 16;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int rs = it & 1;
            const int stage = (q.stages == 2) ? (it & 1) : 0;
            const uint32_t par = ((q.stages == 2) ? (it >> 1) : it) & 1;
            const int row = t / q.tiles_per_row;
            const int ct = t - row * q.tiles_per_row;
            const long long ipA = (q.first_col + static_cast<long long>(ct) * TN - q.pmax) * TB - q.org;
            const bool bulk = tile_bulk(ipA);
            float* raw = reinterpret_cast<float*>(smem_raw + raw0_off + static_cast<uint32_t>(rs) * raw_bytes);
            mbar_wait(BAR(RAW_FULL, rs), (it >> 1) & 1);
            if (!bulk) {                                   // zero / held / extended samples, causal order
                const float* __restrict__ xr = p.x + static_cast<long long>(row) * p.ld_x;
                for (int i = tl; i < len; i += kLoadWarps * 32)
                    raw[i] = tload(p, xr, (p.dir > 0) ? (ipA + i) : (p.n_v - 1 - (ipA + i)));
                asm volatile("bar.sync 1, %0;" ::"n"(kLoadWarps * 32) : "memory");
            }
            mbar_wait(BAR(SLAB_EMPTY, stage), par ^ 1u);
            const uint32_t sbase = slab0 + static_cast<uint32_t>(stage) * stage_bytes;
            const bool rev = bulk && p.dir < 0;            // bulk tiles of an anticausal pass sit in ascending memory order
            for (int item = tl; item < nitems; item += kLoadWarps * 32) {
                const int cidx = item >> 4, sc = item & 15;
                const int o = cidx * TB + sc * 8;
                const float* src = rev ? (raw + (len - 8 - o)) : (raw + o);
                const float4 a = *reinterpret_cast<const float4*>(src);
                const float4 b = *reinterpret_cast<const float4*>(src + 4);
                float v[8];
                v[0] = rev ? b.w : a.x; v[1] = rev ? b.z : a.y; v[2] = rev ? b.y : a.z; v[3] = rev ? b.x : a.w;
                v[4] = rev ? a.w : b.x; v[5] = rev ? a.z : b.y; v[6] = rev ? a.y : b.z; v[7] = rev ? a.x : b.w;
                split_store(v, q.nver, sbase + (static_cast<uint32_t>(sc) * q.slab_cols + cidx) * 16u, ver_bytes);
            }
            fence_proxy_async_smem();                      // my generic-proxy writes -> visible to the MMA's async reads
            mbar_arrive(BAR(SLAB_FULL, stage));
            mbar_arrive(BAR(RAW_EMPTY, rs));               // raw[rs] may be overwritten by the next bulk copy
        }
    } else if (warp >= kEpiWarps && warp < kEpiWarps + kLoadWarps) {
        // ===== LOADER (long filters): stage the tile's N + P columns as split-BF16 core-matrix rows ==========
        // A thread owns items tl, tl + 256, ... (item = (column, 8-sample chunk)); its share of the NEXT
        // tile is fetched into registers right after the current one is stored, so the HBM latency is
        // spent while the tensor core works and only the convert + STS.128 phase waits on slab_empty.
        constexpr int RMAX = ((TN + 32) * 16 + kLoadWarps * 32 - 1) / (kLoadWarps * 32);      // pmax <= 32 on this path
        const int tl = tid - kEpiWarps * 32;
        const int nitems = (TN + q.pmax) * 16;
        float4 raw[RMAX][2];
        uint32_t fastmask = 0;
        auto fetch = [&](int t) {
            const int row = t / q.tiles_per_row;
            const int ct = t - row * q.tiles_per_row;
            const long long j0 = q.first_col + static_cast<long long>(ct) * TN;
            const float* __restrict__ xr = p.x + static_cast<long long>(row) * p.ld_x;
            fastmask = 0;
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int item = tl + r * (kLoadWarps * 32);
                if (item < nitems) {
                    const int cidx = item >> 4, sc = item & 15;
                    const long long ip0 = (j0 - q.pmax + cidx) * TB + sc * 8 - q.org;   // causal index of the chunk's first sample
                    // virtual index of the chunk's lowest address: ip0 (causal) or n_v-1-ip0-7 (anticausal)
                    const long long lo = (p.dir > 0) ? ip0 : (p.n_v - 8 - ip0);
                    if (q.fast_ok && lo >= q.fast_lo && lo + 8 <= q.fast_hi) {
                        raw[r][0] = *reinterpret_cast<const float4*>(xr + lo + p.in_off);
                        raw[r][1] = *reinterpret_cast<const float4*>(xr + lo + p.in_off + 4);
                        fastmask |= 1u << r;
                    } else {                                                   // edge chunk: causal order, rule by rule
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = tload(p, xr, (p.dir > 0) ? (ip0 + e) : (p.n_v - 1 - ip0 - e));
                        raw[r][0] = make_float4(v[0], v[1], v[2], v[3]);
                        raw[r][1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                }
            }
        };
        int it = 0;
        if (static_cast<int>(blockIdx.x) < ntiles) fetch(blockIdx.x);
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int stage = (q.stages == 2) ? (it & 1) : 0;
            const uint32_t par = ((q.stages == 2) ? (it >> 1) : it) & 1;
            mbar_wait(BAR(SLAB_EMPTY, stage), par ^ 1u);
            const uint32_t sbase = slab0 + static_cast<uint32_t>(stage) * stage_bytes;
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int item = tl + r * (kLoadWarps * 32);
                if (item < nitems) {
                    const int cidx = item >> 4, sc = item & 15;
                    const float4 a = raw[r][0], b = raw[r][1];
                    const bool rev = (p.dir < 0) && ((fastmask >> r) & 1u);   // fast anticausal chunks were loaded ascending
                    float v[8];
                    v[0] = rev ? b.w : a.x; v[1] = rev ? b.z : a.y; v[2] = rev ? b.y : a.z; v[3] = rev ? b.x : a.w;
                    v[4] = rev ? a.w : b.x; v[5] = rev ? a.z : b.y; v[6] = rev ? a.y : b.z; v[7] = rev ? a.x : b.w;
                    split_store(v, q.nver, sbase + (static_cast<uint32_t>(sc) * q.slab_cols + cidx) * 16u, ver_bytes);
                }
            }
            fence_proxy_async_smem();                      // my generic-proxy writes -> visible to the MMA's async reads
            mbar_arrive(BAR(SLAB_FULL, stage));
            if (t + static_cast<int>(gridDim.x) < ntiles) fetch(t + gridDim.x);
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA ISSUER ===================================================================================
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int stage = (q.stages == 2) ? (it & 1) : 0;
            const uint32_t spar = ((q.stages == 2) ? (it >> 1) : it) & 1;
            const int acc = it & 1;
            const uint32_t apar = (it >> 1) & 1;
            mbar_wait(BAR(ACC_EMPTY, acc), apar ^ 1u);
            mbar_wait(BAR(SLAB_FULL, stage), spar);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sbase = slab0 + static_cast<uint32_t>(stage) * stage_bytes;
                const uint32_t lbo_b = static_cast<uint32_t>(q.slab_cols) * 16u;
                const uint32_t dcol = tmem_base + static_cast<uint32_t>(acc * TN);
                // only the 14-bit start-address field changes between MMAs: add offsets in 16-byte units
                const uint64_t a_h0 = smem_desc(smem0, 128u, 128u), x_h0 = smem_desc(sbase, lbo_b, 128u);
                const uint64_t hank16 = hank_bytes >> 4, ver16 = ver_bytes >> 4;
                const int terms = q.terms;
                uint32_t accum = 0;
                for (int pb = 0; pb <= q.pmax; ++pb) {
                    const int s_lo = max(0, TB * pb - (q.k - 1));       // first s with a non-zero tap in T_p
                    for (int ks = s_lo >> 4; ks < TB / 16; ++ks) {
                        const uint64_t a_h = a_h0 + static_cast<uint64_t>(8 * (16 * (q.pmax - pb) + 2 * ks));
                        const uint64_t x_h = x_h0 + static_cast<uint64_t>(2 * ks * q.slab_cols + (q.pmax - pb));
                        const uint64_t a_m = a_h + hank16, x_m = x_h + ver16;
                        tc_mma_bf16(dcol, a_h, x_h, kIdesc, accum);     // hi * hi
                        tc_mma_bf16(dcol, a_h, x_m, kIdesc, 1u);        // hi * mid
                        tc_mma_bf16(dcol, a_m, x_h, kIdesc, 1u);        // mid * hi
                        if (terms >= 4) tc_mma_bf16(dcol, a_m, x_m, kIdesc, 1u);
                        if (terms == 6) {
                            tc_mma_bf16(dcol, a_h, x_m + ver16, kIdesc, 1u);     // hi * lo
                            tc_mma_bf16(dcol, a_m + hank16, x_h, kIdesc, 1u);    // lo * hi
                        }
                        accum = 1;
                    }
                }
                tc_commit(BAR(SLAB_EMPTY, stage));         // slab may be refilled once these MMAs retire
                tc_commit(BAR(ACC_FULL, acc));             // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== EPILOGUE: TMEM -> registers -> coalesced row stores ==========================================
        int it = 0;
        const int rr = TB - 1 - (warp * 32 + lane);        // accumulator rows are reversed (Hankel trick)
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t apar = (it >> 1) & 1;
            const int row = t / q.tiles_per_row;
            const int ct = t - row * q.tiles_per_row;
            const long long j0 = q.first_col + static_cast<long long>(ct) * TN;
            float* __restrict__ yr = p.y + static_cast<long long>(row) * p.ld_y + p.out_off;
            mbar_wait(BAR(ACC_FULL, acc), apar);
            tc_fence_after();
            const long long ip_first = j0 * TB - q.org, ip_last = (j0 + TN) * TB - q.org;      // this tile's outputs [first, last)
            const bool interior = ip_first >= q.ip_lo && ip_last <= q.ip_hi;
#pragma unroll 1
            for (int c4 = 0; c4 < TN / 32; ++c4) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(acc * TN + c4 * 32), v);
                const long long ipb = (j0 + c4 * 32) * TB + rr - q.org;
                if (interior) {                            // one base pointer, immediate offsets, no guards
                    float* yb = yr + map_index(p, ipb);
                    if (p.dir > 0) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) yb[c * TB] = __uint_as_float(v[c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) yb[-c * TB] = __uint_as_float(v[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const long long ip = ipb + static_cast<long long>(c) * TB;
                        if (ip >= q.ip_lo && ip < q.ip_hi) yr[map_index(p, ip)] = __uint_as_float(v[c]);
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
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

struct ToepPlan {
    ToepParams q;
    size_t smem_bytes;
};

bool make_plan(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k, ToepPlan* out)
{
    if (k < 1 || k > SCIR_B200_MAX_TAPS) return false;
    ToepParams q{};
    q.p = pass;
    q.k = static_cast<int>(k);
    q.pmax = static_cast<int>((k - 1 + (TB - 1)) / TB);
    if (q.pmax > 32) return false;                         // loader register budget (RMAX) and shared memory
    q.hank_cores = 16 * q.pmax + 31;
    q.slab_cols = (TN + q.pmax) | 1;
    int64_t terms = ctx->opt.toeplitz_terms;
    if (terms != 3 && terms != 4 && terms != 6) terms = 4;
    q.terms = static_cast<int>(terms);
    q.nver = (terms == 6) ? 3 : 2;
    const size_t hank = static_cast<size_t>(q.hank_cores) * 128 * q.nver;
    const size_t stage = static_cast<size_t>(16) * q.slab_cols * 16 * q.nver;
    const size_t budget = static_cast<size_t>(ctx->max_smem_optin) - 1024;        // static smem + slack
    if (hank + stage > budget) return false;
    const size_t raw = static_cast<size_t>(TN + q.pmax) * TB * 4;
    if (ctx->opt.toeplitz_loader != 1 && hank + stage + 2 * raw <= budget) {
        q.raw_stages = 2;                                  // short filters: TMA-prefetched fp32 ring + converter warps
        q.stages = (hank + 2 * stage + 2 * raw <= budget) ? 2 : 1;
    } else {
        q.raw_stages = 0;                                  // long filters: register prefetch (no smem left for a ring)
        q.stages = (hank + 2 * stage <= budget) ? 2 : 1;
    }
    // outputs wanted, in causal index space i'
    q.ip_lo = (pass.dir > 0) ? pass.out_begin : (pass.n_v - pass.out_end);
    q.ip_hi = (pass.dir > 0) ? pass.out_end : (pass.n_v - pass.out_begin);
    // origin shift: chunk addresses are (ip0 + in_off) causal, (n_v - 8 - ip0 + in_off) anticausal, ip0 = 8m - org
    {
        const long long al0 = (pass.dir > 0) ? pass.in_off : -(pass.n_v + pass.in_off);
        q.org = ((al0 % 4) + 4) % 4;
    }
    const long long tile_len = static_cast<long long>(TB) * TN;
    const long long first_tile = (q.ip_lo + q.org) / tile_len;          // ip + org >= 0 is the block-grid coordinate
    q.first_col = first_tile * TN;
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
    out->smem_bytes = hank + static_cast<size_t>(q.stages) * stage + static_cast<size_t>(q.raw_stages) * raw;
    return true;
}

}  // namespace

bool toeplitz_supported(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k)
{
    ToepPlan plan;
    return pass.batch > 0 && pass.out_end > pass.out_begin && make_plan(ctx, pass, k, &plan);
}

int launch_fir_toeplitz(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k)
{
    if (pass.batch <= 0 || pass.out_end <= pass.out_begin) return SCIR_B200_OK;
    ToepPlan plan;
    if (!make_plan(ctx, pass, k, &plan))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "tcgen05 Toeplitz path cannot serve k=%lld", (long long)k);
    SCIR_TRY(ctx_bind(ctx));
    thread_local ToepTaps* tl = nullptr;
    if (!tl) tl = new ToepTaps();
    for (int i = 0; i < SCIR_B200_MAX_TAPS; ++i) tl->c[i] = (i < k) ? c[i] : 0.f;
    static thread_local size_t configured[16] = {};
    const int d = ctx->device & 15;
    if (configured[d] < plan.smem_bytes) {
        SCIR_CUDA(cudaFuncSetAttribute(fir_toeplitz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(plan.smem_bytes)),
                  "cudaFuncSetAttribute(fir_toeplitz_kernel)");
        configured[d] = plan.smem_bytes;
    }
    const int grid = std::min(plan.q.total_tiles, ctx->sm_count);
    fir_toeplitz_kernel<<<grid, kToepThreads, plan.smem_bytes, ctx->stream>>>(plan.q, *tl);
    SCIR_CUDA(cudaGetLastError(), "fir_toeplitz_kernel launch");
    ctx->launches++;
    ctx->toeplitz_launches++;
    return SCIR_B200_OK;
}

}  // namespace scir_b200
