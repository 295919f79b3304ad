// upfirdn.cu -- polyphase up-FIR-down for sm_100a (SciPy semantics, mode='constant').
//
// Spec: scipy/signal/_upfirdn_apply.pyx:421-481.  Closed form evaluated per output m:
//     t = (m*down) % up,  q = (m*down) / up,   y[m] = sum_{i>=0} h[t + i*up] * x[q - i],
//     0 <= q-i < n_in,  t + i*up < len_h
// i.e. the zero-stuffed samples the reference's legacy resampler multiplies by
// (crates/scir-signal/src/lib.rs:348-352) are never touched.
#include "common.cuh"
#include "ext_modes.cuh"

#include <algorithm>

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

// ---- tiled kernel for any rate -------------------------------------------------------------------------------
// The template kernels (upfirdn_poly.cu) cover ten common rates; everything else (160/147, 5/4, ...) used to fall
// to one thread per output with global loads: 0.36 TB/s.  Here a CTA stages the input span of J*L consecutive
// outputs and the phase-transposed taps ht[t][i] = h[t + i*up] in shared memory; L is a multiple of `up`, so the J
// outputs m, m+L, ..., m+(J-1)L of one thread slot share their phase: one tap load feeds J FMAs, and their samples
// sit exactly dq = L*down/up apart.
struct GenTiledParams {
    const float* x;
    float* y;
    const float* ht;              // [up][hpp_pad], device memory
    long long ld_x, ld_y, n_in;
    long long m_begin, m_end;     // outputs [m_begin, m_end) land at y[row, m - m_begin]
    long long tiles_per_row;
    long long up, down;
    int hpp, hpp_pad;             // taps per phase, padded row pitch (odd: bank spread)
    int L, J, dq;                 // slot count, outputs per slot, sample distance between a slot's outputs
    int in_cap;                   // floats reserved for the input span
    ExtSpec ext;
};

template <int J>
__global__ void __launch_bounds__(256) upfirdn_tiled_kernel(const __grid_constant__ GenTiledParams q)
{
    extern __shared__ __align__(16) float sm[];
    float* ts = sm;                                        // transposed taps
    float* xs = sm + q.up * q.hpp_pad;                     // input span
    const long long row = blockIdx.x / q.tiles_per_row;
    const long long tile = blockIdx.x - row * q.tiles_per_row;
    const long long m0 = q.m_begin + tile * (static_cast<long long>(J) * q.L);
    const long long m_last = m0 + static_cast<long long>(J) * q.L - 1;
    const long long q_first = (m0 * q.down) / q.up - (q.hpp - 1);
    const int span = static_cast<int>((m_last * q.down) / q.up - q_first + 1);
    const float* __restrict__ xr = q.x + row * q.ld_x;

    for (int i = threadIdx.x; i < q.up * q.hpp_pad; i += blockDim.x) ts[i] = q.ht[i];
    if (q_first >= 0 && q_first + span <= q.n_in) {
        for (int s2 = threadIdx.x; s2 < span; s2 += blockDim.x) xs[s2] = xr[q_first + s2];
    } else {
        const bool zpad = (q.ext.mode == SCIR_B200_EXT_CONSTANT && q.ext.cval == 0.f);
        for (int s2 = threadIdx.x; s2 < span; s2 += blockDim.x) {
            const long long xi = q_first + s2;
            xs[s2] = (xi >= 0 && xi < q.n_in) ? xr[xi] : (zpad ? 0.f : upfirdn_sample(xr, xi, q.n_in, q.ext));
        }
    }
    __syncthreads();

    float* __restrict__ yr = q.y + row * q.ld_y;
    for (int u = threadIdx.x; u < q.L; u += blockDim.x) {
        const long long m = m0 + u;
        const long long md = m * q.down;
        const int t = static_cast<int>(md % q.up);
        const int qrel = static_cast<int>(md / q.up - q_first);              // newest sample of output m in xs
        const float* trow = ts + t * q.hpp_pad;
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        for (int i = q.hpp - 1; i >= 0; --i) {                               // oldest sample first (pyx:451-453)
            const float c = trow[i];
            const float* xp = xs + (qrel - i);
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = fmaf(c, xp[j * q.dq], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const long long mj = m + static_cast<long long>(j) * q.L;
            if (mj < q.m_end) yr[mj - q.m_begin] = acc[j];
        }
    }
}

// returns SCIR_B200_OK with *handled=false when the filter / rate does not fit shared memory
static int launch_upfirdn_tiled(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                                const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y,
                                int64_t ld_y, int64_t m_begin, int64_t m_count, ExtSpec ext, bool* handled)
{
    *handled = false;
    const int64_t hpp = (len_h + up - 1) / up;
    const int64_t hpp_pad = hpp | 1;
    const int64_t tap_floats = up * hpp_pad;
    const size_t budget = static_cast<size_t>(ctx->max_smem_optin) - 1024;
    if (up > 4096 || hpp > 8192 || static_cast<size_t>(tap_floats) * 4 > budget / 2) return SCIR_B200_OK;
    // slots: a multiple of `up` close to the 256 threads
    int64_t kf = std::max<int64_t>(1, 256 / up), kc = (256 + up - 1) / up;
    const double eff_f = static_cast<double>(kf * up) / (((kf * up + 255) / 256) * 256.0);
    const double eff_c = static_cast<double>(kc * up) / (((kc * up + 255) / 256) * 256.0);
    const int64_t L = ((eff_c > eff_f) ? kc : kf) * up;
    const int64_t dq = L / up * down;
    int J = 8;
    auto in_floats = [&](int j) { return static_cast<int64_t>(j) * L * down / up + hpp + 8; };
    while (J > 1 && (static_cast<size_t>(tap_floats + in_floats(J)) * 4 > std::min<size_t>(budget, 100 * 1024))) J >>= 1;
    if (static_cast<size_t>(tap_floats + in_floats(J)) * 4 > budget) return SCIR_B200_OK;
    const int64_t per_tile = static_cast<int64_t>(J) * L;
    const int64_t tiles = (m_count + per_tile - 1) / per_tile;
    if (tiles * batch > 0x7fffffffLL) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));

    // transposed, padded taps in a ctx-owned device buffer (re-uploaded only when they change)
    std::vector<float> ht(static_cast<size_t>(tap_floats), 0.f);
    for (int64_t t = 0; t < up; ++t)
        for (int64_t i = 0; i < hpp; ++i) {
            const int64_t src = t + i * up;
            if (src < len_h) ht[static_cast<size_t>(t * hpp_pad + i)] = h[src];
        }
    SCIR_TRY(ctx_scratch(ctx, ctx->gen_taps, ht.size() * sizeof(float)));
    if (ctx->gen_taps_host != ht) {
        ctx->gen_taps_host = ht;
        SCIR_CUDA(cudaMemcpyAsync(ctx->gen_taps.ptr, ctx->gen_taps_host.data(), ht.size() * sizeof(float), cudaMemcpyHostToDevice,
                                  ctx->stream),
                  "cudaMemcpyAsync(polyphase taps)");
    }

    GenTiledParams q{};
    q.x = d_x; q.y = d_y; q.ht = static_cast<const float*>(ctx->gen_taps.ptr);
    q.ld_x = ld_x; q.ld_y = ld_y; q.n_in = n_in;
    q.m_begin = m_begin; q.m_end = m_begin + m_count; q.tiles_per_row = tiles;
    q.up = up; q.down = down; q.hpp = static_cast<int>(hpp); q.hpp_pad = static_cast<int>(hpp_pad);
    q.L = static_cast<int>(L); q.J = J; q.dq = static_cast<int>(dq); q.in_cap = static_cast<int>(in_floats(J));
    q.ext = ext;
    const size_t smem = static_cast<size_t>(tap_floats + in_floats(J)) * 4;
    auto launch = [&](auto kern) -> int {
        static thread_local size_t configured[16][4] = {};
        const int d = ctx->device & 15, slot = (J == 8) ? 0 : (J == 4) ? 1 : (J == 2) ? 2 : 3;
        if (smem > 48 * 1024 && configured[d][slot] < smem) {
            SCIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                      "cudaFuncSetAttribute(upfirdn_tiled_kernel)");
            configured[d][slot] = smem;
        }
        kern<<<static_cast<unsigned>(tiles * batch), 256, smem, ctx->stream>>>(q);
        SCIR_CUDA(cudaGetLastError(), "upfirdn_tiled_kernel launch");
        return SCIR_B200_OK;
    };
    int rc;
    if (J == 8) rc = launch(upfirdn_tiled_kernel<8>);
    else if (J == 4) rc = launch(upfirdn_tiled_kernel<4>);
    else if (J == 2) rc = launch(upfirdn_tiled_kernel<2>);
    else rc = launch(upfirdn_tiled_kernel<1>);
    SCIR_TRY(rc);
    ctx->launches++;
    ctx->gen_tiled_launches++;
    *handled = true;
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
    if (ctx->opt.upfirdn_variant != 1 && ctx->opt.upfirdn_variant != 2) {      // 1, 2: force the one-thread-per-output kernel
        bool handled = false;
        SCIR_TRY(launch_upfirdn_tiled(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count, ext, &handled));
        if (handled) return SCIR_B200_OK;
    }
    return launch_upfirdn_generic(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count, ext);
}

}  // namespace scir_b200
