// microbench.cu -- what THIS box sustains, measured in the same process as the bench so the
// roofline denominators carry the same clocks: FP32 FFMA issue peak in the exact instruction form
// the FIR kernel uses (FFMA R, R.reuse, UR, R), and a plain float4 copy for HBM read+write.
#include "common.cuh"

#include <string.h>
#include <algorithm>

namespace scir_b200 {

struct MbTaps {
    float c[64];
};

template <int R>
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, const __grid_constant__ MbTaps taps)
{
    float acc[R];
    float x[4];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = static_cast<float>(r);
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = 1e-3f * static_cast<float>(threadIdx.x + e);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 64; ++t) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(taps.c[t], x[(t + r) & 3], acc[r]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) s += acc[r];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // defeat DCE, never true
}

// Packed form: FFMA2 Racc.F32x2, Rx.F32 (broadcast), URtaps.F32x2, Racc.F32x2 -- two FMAs per issue slot.
// MIX extra integer instructions per 8 FFMA2 probe how many non-FMA issue slots ride along for free.
__device__ __forceinline__ void fma2(unsigned long long& acc, float x, float t0, float t1)
{
    unsigned long long xx, tt;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(tt) : "f"(t0), "f"(t1));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(tt));
}

template <int R2, int MIX>
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float* out, int iters, const __grid_constant__ MbTaps taps)
{
    unsigned long long acc[R2];
    float x[4];
    unsigned junk = threadIdx.x;
#pragma unroll
    for (int r = 0; r < R2; ++r) acc[r] = 0ull;
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = 1e-3f * static_cast<float>(threadIdx.x + e);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 64; t += 2) {
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                fma2(acc[r], x[(t + r) & 3], taps.c[t], taps.c[t + 1]);
                if (MIX > 0 && (r % (8 / MIX)) == 0) junk = junk * 3u + 1u;       // one IMAD per 8/MIX FFMA2
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R2; ++r) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[r]));
        s += lo + hi;
    }
    if (s == 123.456f || junk == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // defeat DCE
}

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4)
{
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i < n4; i += stride) dst[i] = src[i];
}

}  // namespace scir_b200

using namespace scir_b200;

extern "C" {

int scir_b200_microbench_ffma(scir_b200_ctx* ctx, int iters, double* tflops)
{
    SCIR_ENTER(ctx);
    if (!tflops || iters < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "bad microbench arguments");
    SCIR_TRY(ctx_bind(ctx));
    constexpr int R = 20;
    const int blocks = ctx->sm_count * 8;
    float* d_out = nullptr;
    SCIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_out), static_cast<size_t>(blocks) * 256 * 4), "cudaMalloc");
    MbTaps taps;
    for (int i = 0; i < 64; ++i) taps.c[i] = 1e-4f * static_cast<float>(i + 1);
    cudaEvent_t e0, e1;
    SCIR_CUDA(cudaEventCreate(&e0), "cudaEventCreate");
    SCIR_CUDA(cudaEventCreate(&e1), "cudaEventCreate");
    ffma_peak_kernel<R><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps);       // warm-up
    SCIR_CUDA(cudaEventRecord(e0, ctx->stream), "cudaEventRecord");
    ffma_peak_kernel<R><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps);
    SCIR_CUDA(cudaEventRecord(e1, ctx->stream), "cudaEventRecord");
    SCIR_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
    ctx->launches += 2;
    float ms = 0.f;
    SCIR_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
    const double flops = 2.0 * blocks * 256.0 * iters * 64.0 * R;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return SCIR_B200_OK;
}

// mix = 0: pure FFMA2 stream; mix = 1, 2, 4, 8: that many extra integer instructions per 8 FFMA2
int scir_b200_microbench_ffma2(scir_b200_ctx* ctx, int iters, int mix, double* tflops)
{
    SCIR_ENTER(ctx);
    if (!tflops || iters < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "bad microbench arguments");
    SCIR_TRY(ctx_bind(ctx));
    constexpr int R2 = 10;
    const int blocks = ctx->sm_count * 8;
    float* d_out = nullptr;
    SCIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_out), static_cast<size_t>(blocks) * 256 * 4), "cudaMalloc");
    MbTaps taps;
    for (int i = 0; i < 64; ++i) taps.c[i] = 1e-4f * static_cast<float>(i + 1);
    cudaEvent_t e0, e1;
    SCIR_CUDA(cudaEventCreate(&e0), "cudaEventCreate");
    SCIR_CUDA(cudaEventCreate(&e1), "cudaEventCreate");
    auto launch = [&]() {
        switch (mix) {
            case 1: ffma2_peak_kernel<R2, 1><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps); break;
            case 2: ffma2_peak_kernel<R2, 2><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps); break;
            case 4: ffma2_peak_kernel<R2, 4><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps); break;
            case 8: ffma2_peak_kernel<R2, 8><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps); break;
            default: ffma2_peak_kernel<R2, 0><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, taps); break;
        }
    };
    launch();                                                                          // warm-up
    SCIR_CUDA(cudaEventRecord(e0, ctx->stream), "cudaEventRecord");
    launch();
    SCIR_CUDA(cudaEventRecord(e1, ctx->stream), "cudaEventRecord");
    SCIR_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
    ctx->launches += 2;
    float ms = 0.f;
    SCIR_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
    const double flops = 2.0 * blocks * 256.0 * iters * 32.0 * (2 * R2);      // 32 tap pairs x R2 FFMA2 x 2 FMAs
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return SCIR_B200_OK;
}

int scir_b200_microbench_copy(scir_b200_ctx* ctx, size_t bytes, int iters, double* gbps)
{
    SCIR_ENTER(ctx);
    if (!gbps || iters < 1 || bytes < 16) return set_error(SCIR_B200_ERR_INVALID_ARG, "bad microbench arguments");
    SCIR_TRY(ctx_bind(ctx));
    const size_t n4 = bytes / 16;
    float4 *a = nullptr, *b = nullptr;
    SCIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&a), n4 * 16), "cudaMalloc");
    SCIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&b), n4 * 16), "cudaMalloc");
    SCIR_CUDA(cudaMemsetAsync(a, 0, n4 * 16, ctx->stream), "cudaMemsetAsync");
    cudaEvent_t e0, e1;
    SCIR_CUDA(cudaEventCreate(&e0), "cudaEventCreate");
    SCIR_CUDA(cudaEventCreate(&e1), "cudaEventCreate");
    const int blocks = ctx->sm_count * 16;
    copy_kernel<<<blocks, 512, 0, ctx->stream>>>(a, b, n4);                          // warm-up
    double best = 0.0;
    for (int it = 0; it < iters; ++it) {
        SCIR_CUDA(cudaEventRecord(e0, ctx->stream), "cudaEventRecord");
        copy_kernel<<<blocks, 512, 0, ctx->stream>>>(a, b, n4);
        SCIR_CUDA(cudaEventRecord(e1, ctx->stream), "cudaEventRecord");
        SCIR_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
        float ms = 0.f;
        SCIR_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
        const double g = 2.0 * n4 * 16.0 / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    ctx->launches += static_cast<uint64_t>(iters) + 1;
    *gbps = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return SCIR_B200_OK;
}

// What the PCIe link of THIS box moves with plain pinned copies, as the ceiling for the *_host entry points: `bytes`
// each way, H2D alone, D2H alone, and both at once on two streams (the *_host ring's steady state).  Best of `iters`.
int scir_b200_microbench_pcie(scir_b200_ctx* ctx, size_t bytes, int iters, double* h2d_gbs, double* d2h_gbs,
                              double* duplex_each_gbs)
{
    SCIR_ENTER(ctx);
    if (!h2d_gbs || !d2h_gbs || !duplex_each_gbs || iters < 1 || bytes < 4096)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "bad microbench arguments");
    void *ha = nullptr, *hb = nullptr, *da = nullptr, *db = nullptr;
    cudaStream_t s1 = nullptr, s2 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    auto cleanup = [&]() {
        if (ha) cudaFreeHost(ha);
        if (hb) cudaFreeHost(hb);
        if (da) cudaFree(da);
        if (db) cudaFree(db);
        if (s1) cudaStreamDestroy(s1);
        if (s2) cudaStreamDestroy(s2);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (e2) cudaEventDestroy(e2);
    };
    auto body = [&]() -> int {
        SCIR_CUDA(cudaHostAlloc(&ha, bytes, cudaHostAllocPortable), "cudaHostAlloc");
        SCIR_CUDA(cudaHostAlloc(&hb, bytes, cudaHostAllocPortable), "cudaHostAlloc");
        memset(ha, 1, bytes);
        memset(hb, 2, bytes);
        SCIR_CUDA(cudaMalloc(&da, bytes), "cudaMalloc");
        SCIR_CUDA(cudaMalloc(&db, bytes), "cudaMalloc");
        SCIR_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking), "cudaStreamCreate");
        SCIR_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking), "cudaStreamCreate");
        SCIR_CUDA(cudaEventCreate(&e0), "cudaEventCreate");
        SCIR_CUDA(cudaEventCreate(&e1), "cudaEventCreate");
        SCIR_CUDA(cudaEventCreate(&e2), "cudaEventCreate");
        SCIR_CUDA(cudaMemcpyAsync(da, ha, bytes, cudaMemcpyHostToDevice, s1), "warm-up H2D");
        SCIR_CUDA(cudaMemcpyAsync(hb, db, bytes, cudaMemcpyDeviceToHost, s2), "warm-up D2H");
        SCIR_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        double best_h = 0.0, best_d = 0.0, best_x = 0.0;
        for (int it = 0; it < iters; ++it) {
            float ms = 0.f;
            SCIR_CUDA(cudaEventRecord(e0, s1), "cudaEventRecord");
            SCIR_CUDA(cudaMemcpyAsync(da, ha, bytes, cudaMemcpyHostToDevice, s1), "H2D");
            SCIR_CUDA(cudaEventRecord(e1, s1), "cudaEventRecord");
            SCIR_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
            SCIR_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
            best_h = std::max(best_h, bytes / (ms * 1e-3) / 1e9);
            SCIR_CUDA(cudaEventRecord(e0, s2), "cudaEventRecord");
            SCIR_CUDA(cudaMemcpyAsync(hb, db, bytes, cudaMemcpyDeviceToHost, s2), "D2H");
            SCIR_CUDA(cudaEventRecord(e1, s2), "cudaEventRecord");
            SCIR_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
            SCIR_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
            best_d = std::max(best_d, bytes / (ms * 1e-3) / 1e9);
            // both directions at once: s2 starts when s1 starts; the slower stream's end closes the interval
            SCIR_CUDA(cudaEventRecord(e0, s1), "cudaEventRecord");
            SCIR_CUDA(cudaStreamWaitEvent(s2, e0, 0), "cudaStreamWaitEvent");
            SCIR_CUDA(cudaMemcpyAsync(da, ha, bytes, cudaMemcpyHostToDevice, s1), "H2D");
            SCIR_CUDA(cudaMemcpyAsync(hb, db, bytes, cudaMemcpyDeviceToHost, s2), "D2H");
            SCIR_CUDA(cudaEventRecord(e2, s2), "cudaEventRecord");
            SCIR_CUDA(cudaStreamWaitEvent(s1, e2, 0), "cudaStreamWaitEvent");
            SCIR_CUDA(cudaEventRecord(e1, s1), "cudaEventRecord");
            SCIR_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
            SCIR_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
            best_x = std::max(best_x, bytes / (ms * 1e-3) / 1e9);
        }
        *h2d_gbs = best_h;
        *d2h_gbs = best_d;
        *duplex_each_gbs = best_x;
        return SCIR_B200_OK;
    };
    const int rc = body();
    cleanup();
    return rc;
}

}  // extern "C"
