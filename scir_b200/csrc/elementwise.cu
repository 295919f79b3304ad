// elementwise.cu -- the three f32 elementwise ops of scir-gpu's DeviceArray on device-resident data.
//
// Replaces add_vec_f32_cuda / add_scalar_f32_cuda / mul_scalar_f32_cuda (crates/scir-gpu/src/lib.rs:840-1034)
// and their PTX entries (:640-726), which allocate, copy in, launch one thread per element on the NULL stream,
// synchronise and copy out on every call (and never ran: the PTX module does not assemble).  Here
// the arrays stay in HBM between ops, so the kernel is pure streaming work: 8 B (unary) or
// 12 B (binary) per element, HBM-bound.  128-bit loads/stores, 4 independent float4 per thread per trip
// (memory-level parallelism), grid = a multiple of the SM count, scalar head/tail for unaligned views.
// Results are bit-identical to the reference's CPU loops (:206-255): one IEEE add or mul per element, no FMA.
#include "common.cuh"

#include <algorithm>

namespace scir_b200 {

enum { EW_ADD_SCALAR = 0, EW_MUL_SCALAR = 1, EW_ADD = 2 };

template <int OP>
__device__ __forceinline__ float ew_apply(float a, float b, float alpha)
{
    if (OP == EW_ADD_SCALAR) return __fadd_rn(a, alpha);
    if (OP == EW_MUL_SCALAR) return __fmul_rn(a, alpha);
    return __fadd_rn(a, b);
}

template <int OP>
__global__ void __launch_bounds__(256) elementwise_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          float alpha, float* __restrict__ y, long long n, int vec_ok)
{
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    if (vec_ok) {
        const long long n4 = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        float4* y4 = reinterpret_cast<float4*>(y);
        constexpr int U = 4;
        long long i = tid;
        for (; i + (U - 1) * nthreads < n4; i += U * nthreads) {
            float4 va[U], vb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                va[u] = __ldcs(a4 + i + u * nthreads);                 // streaming: touched once
                if (OP == EW_ADD) vb[u] = __ldcs(b4 + i + u * nthreads);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float4 r;
                r.x = ew_apply<OP>(va[u].x, vb[u].x, alpha);
                r.y = ew_apply<OP>(va[u].y, vb[u].y, alpha);
                r.z = ew_apply<OP>(va[u].z, vb[u].z, alpha);
                r.w = ew_apply<OP>(va[u].w, vb[u].w, alpha);
                __stcs(y4 + i + u * nthreads, r);
            }
        }
        for (; i < n4; i += nthreads) {
            const float4 va = a4[i];
            float4 vb = va;
            if (OP == EW_ADD) vb = b4[i];
            float4 r;
            r.x = ew_apply<OP>(va.x, vb.x, alpha);
            r.y = ew_apply<OP>(va.y, vb.y, alpha);
            r.z = ew_apply<OP>(va.z, vb.z, alpha);
            r.w = ew_apply<OP>(va.w, vb.w, alpha);
            y4[i] = r;
        }
        for (long long j = (n4 << 2) + tid; j < n; j += nthreads) y[j] = ew_apply<OP>(a[j], (OP == EW_ADD) ? b[j] : 0.f, alpha);
    } else {
        for (long long j = tid; j < n; j += nthreads) y[j] = ew_apply<OP>(a[j], (OP == EW_ADD) ? b[j] : 0.f, alpha);
    }
}

int launch_elementwise(scir_b200_ctx* ctx, int op, const float* d_a, const float* d_b, float alpha, float* d_y, int64_t n)
{
    if (n == 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    const int vec_ok = aligned16(d_a) && aligned16(d_y) && (op != EW_ADD || aligned16(d_b));
    const long long per_cta = 256LL * 4 * 4;                                     // elements one CTA covers per trip
    long long grid = (n + per_cta - 1) / per_cta;
    grid = std::max<long long>(1, std::min<long long>(grid, static_cast<long long>(ctx->sm_count) * 16));
    const unsigned g = static_cast<unsigned>(grid);
    switch (op) {
        case EW_ADD_SCALAR: elementwise_kernel<EW_ADD_SCALAR><<<g, 256, 0, ctx->stream>>>(d_a, d_b, alpha, d_y, n, vec_ok); break;
        case EW_MUL_SCALAR: elementwise_kernel<EW_MUL_SCALAR><<<g, 256, 0, ctx->stream>>>(d_a, d_b, alpha, d_y, n, vec_ok); break;
        default: elementwise_kernel<EW_ADD><<<g, 256, 0, ctx->stream>>>(d_a, d_b, alpha, d_y, n, vec_ok); break;
    }
    SCIR_CUDA(cudaGetLastError(), "elementwise_kernel launch");
    ctx->launches++;
    return SCIR_B200_OK;
}

// ---- per-row statistics and per-row offsets: resample_poly's padtype = mean / minimum / maximum --------------
// (scipy/signal/_signaltools.py:3921-3957: background = stat(x, axis), x - background goes through a zero-padded
// upfirdn, background is added back).  One CTA per row, f32 pairwise-ish tree like numpy's f32 mean.
enum { STAT_MEAN = 0, STAT_MIN = 1, STAT_MAX = 2 };

template <int STAT>
__global__ void __launch_bounds__(512) row_stat_kernel(const float* __restrict__ x, long long ld_x, long long n, float* __restrict__ out)
{
    const float* xr = x + static_cast<long long>(blockIdx.x) * ld_x;
    float v = (STAT == STAT_MEAN) ? 0.f : xr[0];
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float t = xr[i];
        v = (STAT == STAT_MEAN) ? (v + t) : (STAT == STAT_MIN ? fminf(v, t) : fmaxf(v, t));
    }
    __shared__ float sh[512];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = 256; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float a = sh[threadIdx.x], b = sh[threadIdx.x + s];
            sh[threadIdx.x] = (STAT == STAT_MEAN) ? (a + b) : (STAT == STAT_MIN ? fminf(a, b) : fmaxf(a, b));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (STAT == STAT_MEAN) ? sh[0] / static_cast<float>(n) : sh[0];
}

// y[row, i] = x[row, i] + sign * bg[row]
__global__ void __launch_bounds__(256) row_offset_kernel(const float* x, long long ld_x, const float* __restrict__ bg, float sign,
                                                         float* y, long long ld_y, long long n)       // y may alias x
{
    const long long row = blockIdx.y;
    const float b = sign * bg[row];
    const float* xr = x + row * ld_x;
    float* yr = y + row * ld_y;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        yr[i] = __fadd_rn(xr[i], b);
}

// ---- per-row median (resample_poly padtype='median'): exact k-th order statistics by radix select -----------
// One CTA per row.  Floats map to unsigned keys that sort like the values; four passes of an 8-bit histogram over
// the elements that still match the selected prefix pin down the k-th smallest key exactly.  Even n: numpy's
// median is the f32 mean of the two middle values (np.mean of two f32 -> (a + b) / 2 in f32), so two selections.
__device__ __forceinline__ uint32_t f32_sort_key(float v)
{
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_sort_key(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ float row_select(const float* __restrict__ xr, long long n, long long kth, unsigned* hist, unsigned* sh_state)
{
    uint32_t prefix = 0, mask = 0;
    long long k = kth;                                     // rank among the elements matching the prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = f32_sort_key(xr[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long acc = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (acc + hist[b] > k) break;
                acc += hist[b];
            }
            sh_state[0] = static_cast<unsigned>(b);
            sh_state[1] = static_cast<unsigned>(k - acc);  // fits: a row has < 2^32 elements per bin (n checked by the host)
        }
        __syncthreads();
        prefix |= sh_state[0] << shift;
        mask |= 255u << shift;
        k = sh_state[1];
        __syncthreads();
    }
    return f32_from_sort_key(prefix);
}

__global__ void __launch_bounds__(1024) row_median_kernel(const float* __restrict__ x, long long ld_x, long long n, float* __restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned state[2];
    const float* xr = x + static_cast<long long>(blockIdx.x) * ld_x;
    const float hi = row_select(xr, n, n / 2, hist, state);
    float med = hi;
    if ((n & 1) == 0) {
        const float lo = row_select(xr, n, n / 2 - 1, hist, state);
        med = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
    }
    if (threadIdx.x == 0) out[blockIdx.x] = med;
}

int launch_row_stat(scir_b200_ctx* ctx, int stat, const float* d_x, int64_t ld_x, int64_t batch, int64_t n, float* d_out)
{
    if (batch == 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    if (batch > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "too many rows");
    const unsigned g = static_cast<unsigned>(batch);
    if (stat == 3) {                                       // median
        if (n >= (1LL << 32)) return set_error(SCIR_B200_ERR_UNSUPPORTED, "median: rows of 2^32 samples or more");
        row_median_kernel<<<g, 1024, 0, ctx->stream>>>(d_x, ld_x, n, d_out);
    } else if (stat == STAT_MEAN) row_stat_kernel<STAT_MEAN><<<g, 512, 0, ctx->stream>>>(d_x, ld_x, n, d_out);
    else if (stat == STAT_MIN) row_stat_kernel<STAT_MIN><<<g, 512, 0, ctx->stream>>>(d_x, ld_x, n, d_out);
    else row_stat_kernel<STAT_MAX><<<g, 512, 0, ctx->stream>>>(d_x, ld_x, n, d_out);
    SCIR_CUDA(cudaGetLastError(), "row_stat_kernel launch");
    ctx->launches++;
    return SCIR_B200_OK;
}

int launch_row_offset(scir_b200_ctx* ctx, const float* d_x, int64_t ld_x, const float* d_bg, float sign, float* d_y,
                      int64_t ld_y, int64_t batch, int64_t n)
{
    if (batch == 0 || n == 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    if (batch > 65535) return set_error(SCIR_B200_ERR_UNSUPPORTED, "row offset: more than 65535 rows per call");
    const unsigned gx = static_cast<unsigned>(std::min<long long>((n + 255) / 256, 64));
    row_offset_kernel<<<dim3(gx, static_cast<unsigned>(batch)), 256, 0, ctx->stream>>>(d_x, ld_x, d_bg, sign, d_y, ld_y, n);
    SCIR_CUDA(cudaGetLastError(), "row_offset_kernel launch");
    ctx->launches++;
    return SCIR_B200_OK;
}

}  // namespace scir_b200
