// fir_f64.cu -- double-precision batched FIR pass on the device.
//
// The reference's fir1d_batched_f64 (crates/scir-gpu/src/lib.rs:1166-1184) is a CPU loop that only its own tests
// use (:1263-1298), as the high-precision twin of the f32 path; SURVEY 8(f).4 lists f64 variants as a follow-up.
// This is that twin for device-resident data: y[b,i] = sum_t taps[k-1-t] * x[b,i-t], zero state, products and sums in
// IEEE f64 (DFMA), accumulated from the newest sample to the oldest like the reference -- generalised to the same
// "virtual sequence" pass the f32 kernels implement (common.cuh: FirPass), so that the f64 filtfilt (f64_routes.cu) is
// two launches of it: causal or anticausal direction, zero or held boundary, odd / even / constant extension.
//
// FP64 on B200 is a narrow pipe, so the kernel is DFMA-bound for anything but the shortest filters and is kept
// simple: a CTA stages one 2048-output tile plus its K-1 halo in shared memory (the loader applies extension and
// boundary), a thread owns 8 consecutive outputs and slides a 15-sample register window over the taps, 8 taps at a
// time: per group 64 DFMA with static register indices, 8 LDS.64, 8 broadcast tap loads and 7 register moves.
#include "common.cuh"

namespace scir_b200 {

namespace {

constexpr int kF64Threads = 256;
constexpr int kF64R = 8;
constexpr int kF64Tile = kF64Threads * kF64R;

// virtual input sequence v[i] of a pass (same rules as fir_direct.cu: vload)
__device__ __forceinline__ double vload64(const FirPass64& p, const double* __restrict__ xr, long long i)
{
    if (i < 0) {
        if (p.bound == BOUND_ZERO) return 0.0;
        i = 0;
    } else if (i >= p.n_v) {
        if (p.bound == BOUND_ZERO) return 0.0;
        i = p.n_v - 1;
    }
    const long long u = i + p.in_off;
    if (p.ext_mode == EXT_NONE) return xr[u];
    const long long last = p.n_x - 1;
    if (u < 0) {
        if (p.ext_mode == EXT_ODD) return __dsub_rn(__dmul_rn(2.0, xr[0]), xr[-u]);      // scipy _arraytools.py:57-107
        if (p.ext_mode == EXT_EVEN) return xr[-u];
        return xr[0];
    }
    if (u > last) {
        if (p.ext_mode == EXT_ODD) return __dsub_rn(__dmul_rn(2.0, xr[last]), xr[2 * last - u]);
        if (p.ext_mode == EXT_EVEN) return xr[2 * last - u];
        return xr[last];
    }
    return xr[u];
}

__global__ void __launch_bounds__(kF64Threads) fir_f64_kernel(const __grid_constant__ FirPass64 p, const double* __restrict__ c, int k,
                                                              long long tiles_per_row)
{
    extern __shared__ double win[];                        // [k - 1 + tile] samples, in the order the taps walk them
    const long long row = blockIdx.x / tiles_per_row;
    const long long tile = blockIdx.x - row * tiles_per_row;
    const long long i0 = p.out_begin + tile * kF64Tile;    // first output (virtual index) of this tile
    const double* xr = p.x + row * p.ld_x;
    const int halo = k - 1;
    const int len = halo + kF64Tile;
    // one pad double per 8 samples: a thread's window starts every 9 doubles, so the 16 lanes of an LDS.64 phase
    // hit 16 different bank pairs (stride 8 would be a 16-way conflict)
    auto pidx = [](int s2) { return s2 + (s2 >> 3); };
    // causal:     win[s] = v[i0 - halo + s]            local output o  <-> virtual index i0 + o
    // anticausal: win[s] = v[i0 + tile - 1 + halo - s] local output o' <-> virtual index i0 + tile - 1 - o'
    // either way output o needs win[halo + o - d] for tap d, so the arithmetic below is the same for both directions.
    for (int s = threadIdx.x; s < len; s += kF64Threads) {
        const long long i = (p.dir > 0) ? (i0 - halo + s) : (i0 + kF64Tile - 1 + halo - s);
        win[pidx(s)] = vload64(p, xr, i);
    }
    __syncthreads();

    // Output o = R*tid + j uses win[halo + o - d].  Taps go in groups of R: a 2R-1 sample register window W
    // (W[i] = win[base - d0 - (R-1) + i]) serves all R x R (output, tap) pairs of a group with static indices, then
    // slides down by R (R-1 register moves and R new samples per R*R DFMA).  Taps beyond k are SKIPPED, not multiplied
    // as zeros: a NaN / Inf sample must reach exactly the k outputs whose window holds it, like the reference loop.
    constexpr int R = kF64R;
    const int base = halo + threadIdx.x * R;
    double acc[R], W[2 * R - 1];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.0;
#pragma unroll
    for (int i = 0; i < 2 * R - 1; ++i) {
        const int idx = base - (R - 1) + i;
        W[i] = (idx >= 0) ? win[pidx(idx)] : 0.0;
    }
    for (int d0 = 0; d0 < k; d0 += R) {
        double cd[R];
#pragma unroll
        for (int dd = 0; dd < R; ++dd) cd[dd] = (d0 + dd < k) ? c[d0 + dd] : 0.0;
#pragma unroll
        for (int dd = 0; dd < R; ++dd) {                   // newest sample first within the group, like the reference
            if (d0 + dd < k) {
#pragma unroll
                for (int j = 0; j < R; ++j) acc[j] = fma(cd[dd], W[R - 1 + j - dd], acc[j]);
            }
        }
#pragma unroll
        for (int i = 2 * R - 2; i >= R; --i) W[i] = W[i - R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int idx = base - (d0 + R) - (R - 1) + i;
            W[i] = (idx >= 0) ? win[pidx(idx)] : 0.0;
        }
    }
    double* yr = p.y + row * p.ld_y;
#pragma unroll
    for (int j = 0; j < kF64R; ++j) {
        const int o = threadIdx.x * kF64R + j;
        const long long i = (p.dir > 0) ? (i0 + o) : (i0 + kF64Tile - 1 - o);
        if (i >= p.out_begin && i < p.out_end) yr[i + p.out_off] = acc[j];
    }
}

}  // namespace

int launch_fir_pass_f64(scir_b200_ctx* ctx, const FirPass64& pass, const double* taps_by_delay, int64_t k)
{
    const long long n_out = pass.out_end - pass.out_begin;
    if (pass.batch == 0 || n_out <= 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    const size_t len = static_cast<size_t>(k - 1 + kF64Tile);
    const size_t smem = (len + len / 8 + 1) * sizeof(double);                    // padded: see pidx in the kernel
    if (smem > static_cast<size_t>(ctx->max_smem_optin))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "f64 FIR: %lld taps need %zu B of shared memory", (long long)k, smem);
    const long long tiles = (n_out + kF64Tile - 1) / kF64Tile;
    if (tiles * pass.batch > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "grid too large");
    double* d_c = nullptr;
    SCIR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_c), static_cast<size_t>(k) * sizeof(double), ctx->stream), "cudaMallocAsync(taps)");
    SCIR_CUDA(cudaMemcpyAsync(d_c, taps_by_delay, static_cast<size_t>(k) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream),
              "cudaMemcpyAsync(taps)");
    static thread_local size_t configured[16] = {};
    const int d = ctx->device & 15;
    if (smem > 48 * 1024 && configured[d] < smem) {
        SCIR_CUDA(cudaFuncSetAttribute(fir_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "cudaFuncSetAttribute(fir_f64_kernel)");
        configured[d] = smem;
    }
    fir_f64_kernel<<<static_cast<unsigned>(tiles * pass.batch), kF64Threads, smem, ctx->stream>>>(pass, d_c, static_cast<int>(k), tiles);
    SCIR_CUDA(cudaGetLastError(), "fir_f64_kernel launch");
    ctx->launches++;
    SCIR_CUDA(cudaFreeAsync(d_c, ctx->stream), "cudaFreeAsync(taps)");
    return SCIR_B200_OK;
}

int launch_fir_f64(scir_b200_ctx* ctx, const double* d_x, int64_t ld_x, const double* taps_by_delay, int64_t k, double* d_y,
                   int64_t ld_y, int64_t batch, int64_t n)
{
    FirPass64 p{};
    p.x = d_x; p.y = d_y; p.ld_x = ld_x; p.ld_y = ld_y; p.batch = batch;
    p.n_x = n; p.n_v = n; p.in_off = 0; p.out_off = 0; p.out_begin = 0; p.out_end = n;
    p.ext_mode = EXT_NONE; p.bound = BOUND_ZERO; p.dir = +1;
    return launch_fir_pass_f64(ctx, p, taps_by_delay, k);
}

}  // namespace scir_b200
