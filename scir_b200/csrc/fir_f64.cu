// fir_f64.cu -- double-precision batched FIR on the device.
//
// The reference's fir1d_batched_f64 (crates/scir-gpu/src/lib.rs:1166-1184) is a CPU loop that only its own tests
// use (:1263-1298), as the high-precision twin of the f32 path; SURVEY 8(f).4 lists f64 variants as a follow-up.
// This is that twin for device-resident data: same definition y[b,i] = sum_t taps[k-1-t] * x[b,i-t], zero state,
// products and sums in IEEE f64 (DFMA), accumulated from the newest sample to the oldest like the reference.
//
// FP64 on B200 is a narrow pipe, so the kernel is DFMA-bound for anything but the shortest filters and is kept
// simple: a CTA stages one 2048-output tile plus its K-1 halo in shared memory (coalesced loads), a thread owns
// 8 consecutive outputs and slides a 15-sample register window over the taps, 8 taps at a time: per group 64 DFMA
// with static register indices, 8 LDS.64, 8 broadcast tap loads and 7 register moves.
#include "common.cuh"

namespace scir_b200 {

namespace {

constexpr int kF64Threads = 256;
constexpr int kF64R = 8;
constexpr int kF64Tile = kF64Threads * kF64R;

__global__ void __launch_bounds__(kF64Threads) fir_f64_kernel(const double* __restrict__ x, long long ld_x, const double* __restrict__ c,
                                                              int k, double* __restrict__ y, long long ld_y, long long n,
                                                              long long tiles_per_row)
{
    extern __shared__ double win[];                        // [k - 1 + tile]: samples i0 - (k-1) .. i0 + tile - 1
    const long long row = blockIdx.x / tiles_per_row;
    const long long tile = blockIdx.x - row * tiles_per_row;
    const long long i0 = tile * kF64Tile;
    const double* xr = x + row * ld_x;
    const int halo = k - 1;
    const int len = halo + kF64Tile;
    // one pad double per 8 samples: a thread's window starts every 9 doubles, so the 16 lanes of an LDS.64 phase
    // hit 16 different bank pairs (stride 8 would be a 16-way conflict)
    auto pidx = [](int s2) { return s2 + (s2 >> 3); };
    for (int s = threadIdx.x; s < len; s += kF64Threads) {
        const long long i = i0 - halo + s;
        win[pidx(s)] = (i >= 0 && i < n) ? xr[i] : 0.0;    // zero initial state, zero beyond the row
    }
    __syncthreads();

    // Output o = R*tid + j uses win[halo + o - d].  Taps go in groups of R: a 2R-1 sample register window W
    // (W[i] = win[base - d0 - (R-1) + i]) serves all R x R (output, tap) pairs of a group with static indices, then
    // slides down by R (R-1 register moves and R new samples per R*R DFMA).  The tail group is masked by zero taps.
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
        for (int dd = 0; dd < R; ++dd)                     // newest sample first within the group, like the reference
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = fma(cd[dd], W[R - 1 + j - dd], acc[j]);
#pragma unroll
        for (int i = 2 * R - 2; i >= R; --i) W[i] = W[i - R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int idx = base - (d0 + R) - (R - 1) + i;
            W[i] = (idx >= 0) ? win[pidx(idx)] : 0.0;
        }
    }
    double* yr = y + row * ld_y;
#pragma unroll
    for (int j = 0; j < kF64R; ++j) {
        const long long i = i0 + threadIdx.x * kF64R + j;
        if (i < n) yr[i] = acc[j];
    }
}

}  // namespace

int launch_fir_f64(scir_b200_ctx* ctx, const double* d_x, int64_t ld_x, const double* taps_by_delay, int64_t k, double* d_y,
                   int64_t ld_y, int64_t batch, int64_t n)
{
    if (batch == 0 || n == 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    const size_t len = static_cast<size_t>(k - 1 + kF64Tile);
    const size_t smem = (len + len / 8 + 1) * sizeof(double);                    // padded: see pidx in the kernel
    if (smem > static_cast<size_t>(ctx->max_smem_optin))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "f64 FIR: %lld taps need %zu B of shared memory", (long long)k, smem);
    const long long tiles = (n + kF64Tile - 1) / kF64Tile;
    if (tiles * batch > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "grid too large");
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
    fir_f64_kernel<<<static_cast<unsigned>(tiles * batch), kF64Threads, smem, ctx->stream>>>(d_x, ld_x, d_c, static_cast<int>(k), d_y, ld_y,
                                                                                              n, tiles);
    SCIR_CUDA(cudaGetLastError(), "fir_f64_kernel launch");
    ctx->launches++;
    SCIR_CUDA(cudaFreeAsync(d_c, ctx->stream), "cudaFreeAsync(taps)");
    return SCIR_B200_OK;
}

}  // namespace scir_b200
