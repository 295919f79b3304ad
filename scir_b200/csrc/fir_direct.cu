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
//
// The same kernel serves lfilter / filtfilt: direction (causal / anticausal), boundary rule
// (zero state or held end value = SciPy's lfilter_zi steady state) and odd/even/constant signal
// extension are properties of the tile LOADER (edge tiles only); interior tiles always take the
// bulk-copy path.  See DESIGN.md section 4.
#include "common.cuh"

namespace scir_b200 {

template <int MAXK>
struct TapsParam {
    float c[MAXK];
};

struct FirTileParams {
    FirPass p;
    long long base0;      // virtual index of the first output of tile 0 ((base0 + in_off) % 4 == 0)
    long long ntiles;     // tiles per row
    int nchunk;           // tap chunks (halo = nchunk * KC)
    int in_vec_ok;        // input rows are 16-B aligned => interior tiles may use the bulk copy
    int out_vec_ok;       // same for the output side
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                              uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_copy_s2g(void* dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(src_smem), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

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

// ---- the tile kernel -------------------------------------------------------------------------------
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

    // ---- R outputs per thread, x-major over the window, taps from uniform registers -----------------
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;

    // causal:     smem index of v[i0 + j] is DP + j; chunk c uses samples t0 + r - d, d in [c*KC,(c+1)*KC)
    // anticausal: smem index of v[i0 + j] is j;      chunk c uses samples t0 + r + d
    const float* wbase = (DIR > 0) ? (smem + DP + tid * R) : (smem + tid * R);
    for (int c = 0; c < q.nchunk; ++c) {
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

    // ---- results back through shared memory, then one bulk store --------------------------------------
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
            bulk_store_wait_read();                        // smem must outlive the read
        }
    } else {
        __syncthreads();
        for (int s = tid; s < TILE; s += NT) {
            const long long i = i0 + s;
            if (i >= p.out_begin && i < p.out_end) yr[i + p.out_off] = smem[s];
        }
    }
}

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

// ---- host side ---------------------------------------------------------------------------------------------
namespace {

constexpr int kR = 20;
constexpr int kNT = 256;
constexpr int kTile = kR * kNT;
constexpr int kSmallK = 256;
constexpr int kBigK = SCIR_B200_MAX_TAPS;

template <int KC, int DIR, int MAXK>
int launch_tile(scir_b200_ctx* ctx, const FirTileParams& q, const float* c, int64_t k, size_t smem_bytes,
                long long grid)
{
    // zero-padded host staging; the launch copies it into the parameter buffer synchronously
    thread_local TapsParam<MAXK>* tl = nullptr;
    if (!tl) tl = new TapsParam<MAXK>();
    for (int i = 0; i < MAXK; ++i) tl->c[i] = (i < k) ? c[i] : 0.f;
    auto kern = fir_tile_kernel<KC, kR, kNT, DIR, MAXK>;
    static thread_local size_t configured[16] = {0};
    if (smem_bytes > 48 * 1024 && configured[ctx->device & 15] < smem_bytes) {
        SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes)),
                  "cudaFuncSetAttribute(fir_tile_kernel)");
        configured[ctx->device & 15] = smem_bytes;
    }
    kern<<<static_cast<unsigned>(grid), kNT, smem_bytes, ctx->stream>>>(q, *tl);
    SCIR_CUDA(cudaGetLastError(), "fir_tile_kernel launch");
    ctx->launches++;
    return SCIR_B200_OK;
}

}  // namespace

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
    const size_t smem_bytes = static_cast<size_t>(DP + kTile) * sizeof(float);
    if (smem_bytes > static_cast<size_t>(ctx->max_smem_optin))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "tile needs %zu B of shared memory", smem_bytes);

    const bool small = (DP <= kSmallK);
    if (pass.dir > 0) {
        if (KC == 32) return launch_tile<32, +1, kSmallK>(ctx, q, c, k, smem_bytes, grid);
        if (small) return launch_tile<64, +1, kSmallK>(ctx, q, c, k, smem_bytes, grid);
        return launch_tile<64, +1, kBigK>(ctx, q, c, k, smem_bytes, grid);
    } else {
        if (KC == 32) return launch_tile<32, -1, kSmallK>(ctx, q, c, k, smem_bytes, grid);
        if (small) return launch_tile<64, -1, kSmallK>(ctx, q, c, k, smem_bytes, grid);
        return launch_tile<64, -1, kBigK>(ctx, q, c, k, smem_bytes, grid);
    }
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

}  // namespace scir_b200
