// upfirdn.cu -- polyphase up-FIR-down for sm_100a (SciPy semantics, mode='constant').
//
// Spec: scipy/signal/_upfirdn_apply.pyx:421-481.  Closed form evaluated per output m:
//     t = (m*down) % up,  q = (m*down) / up,   y[m] = sum_{i>=0} h[t + i*up] * x[q - i],
//     0 <= q-i < n_in,  t + i*up < len_h
// i.e. the zero-stuffed samples the reference's legacy resampler multiplies by
// (crates/scir-signal/src/lib.rs:348-352) are never touched.
#include "common.cuh"
#include "upfirdn_ext.cuh"

namespace scir_b200 {

// ---- generic kernel: any up/down, one thread per output ------------------------------------------------
__global__ void upfirdn_generic_kernel(const float* __restrict__ h, long long len_h, long long up, long long down,
                                       const float* __restrict__ x, long long ld_x, long long batch,
                                       long long n_in, float* __restrict__ y, long long ld_y, long long m_begin,
                                       long long m_count, const ExtSpec ext)
{
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= batch * m_count) return;
    const long long row = gid / m_count;
    const long long j = gid - row * m_count;
    const long long md = (m_begin + j) * down;
    const long long t = md % up;
    const long long q = md / up;
    const float* xr = x + row * ld_x;
    // oldest sample first, like the Cython loop (pyx:451-453); with mode='constant', cval=0 samples outside
    // [0, n_in) are skipped (pyx:446-447), otherwise they take the extension's value (pyx:449-452, :466-471)
    const long long imax = (len_h - 1 - t) / up;    // largest i with t + i*up < len_h
    const bool zpad = (ext.mode == SCIR_B200_EXT_CONSTANT && ext.cval == 0.f);
    float acc = 0.f;
    for (long long i = imax; i >= 0; --i) {
        const long long xi = q - i;
        if (xi >= 0 && xi < n_in) acc = fmaf(h[t + i * up], xr[xi], acc);
        else if (!zpad) acc = fmaf(h[t + i * up], upfirdn_sample(xr, xi, n_in, ext), acc);
    }
    y[row * ld_y + j] = acc;
}

int launch_upfirdn_generic(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                           const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y,
                           int64_t ld_y, int64_t m_begin, int64_t m_count, ExtSpec ext)
{
    SCIR_TRY(ctx_bind(ctx));
    float* d_h = nullptr;
    SCIR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_h), static_cast<size_t>(len_h) * 4, ctx->stream),
              "cudaMallocAsync(h)");
    SCIR_CUDA(cudaMemcpyAsync(d_h, h, static_cast<size_t>(len_h) * 4, cudaMemcpyHostToDevice, ctx->stream),
              "cudaMemcpyAsync(h)");
    const long long total = batch * m_count;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "grid too large");
    upfirdn_generic_kernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(
        d_h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count, ext);
    SCIR_CUDA(cudaGetLastError(), "upfirdn_generic_kernel launch");
    ctx->launches++;
    SCIR_CUDA(cudaFreeAsync(d_h, ctx->stream), "cudaFreeAsync(h)");
    return SCIR_B200_OK;
}

int launch_upfirdn_poly(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                        const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y, int64_t ld_y,
                        int64_t m_begin, int64_t m_count, ExtSpec ext, bool* handled);

int launch_upfirdn(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                   const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y, int64_t ld_y,
                   int64_t m_begin, int64_t m_count, int ext_mode, float cval)
{
    if (ext_mode < SCIR_B200_EXT_CONSTANT || ext_mode > SCIR_B200_EXT_LINE)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown extension mode %d", ext_mode);
    // SciPy divides by (len_x - 1) in these modes (pyx:129-131, :144, :160-163): a single sample has no slope / mirror
    const bool needs2 = ext_mode == SCIR_B200_EXT_REFLECT || ext_mode == SCIR_B200_EXT_SMOOTH ||
                        ext_mode == SCIR_B200_EXT_LINE || ext_mode == SCIR_B200_EXT_ANTIREFLECT;
    if (needs2 && n_in < 2)
        return set_error(SCIR_B200_ERR_SHAPE, "extension mode %d needs at least two samples per row", ext_mode);
    const ExtSpec ext{ext_mode, cval};
    if (ctx->opt.upfirdn_variant != 1) {
        bool handled = false;
        SCIR_TRY(launch_upfirdn_poly(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin,
                                     m_count, ext, &handled));
        if (handled) return SCIR_B200_OK;
    }
    return launch_upfirdn_generic(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count, ext);
}

}  // namespace scir_b200
