// f64_routes.cu -- double-precision twins of the scir-signal FIR routes on the device.
//
// The reference's own `resample_poly` and `filtfilt` take and return Array1<f64> (crates/scir-signal/src/lib.rs:278-291,
// :313-362) and its fixtures are f64 (:638-668), so these routes exist in f64 too: upfirdn with SciPy's nine extension
// modes (scipy/signal/_upfirdn_apply.pyx:110-231, :421-481), resample_poly with every padtype
// (_signaltools.py:3865-3957) and FIR filtfilt (_signaltools.py:4745-4826, plus the reference's zero-state structure).
// Arithmetic is IEEE f64 (DFMA); B200's FP64 pipe is narrow, so these kernels are written for clarity and correct
// edges, not for a roofline: a staged shared-memory tile, taps read through the read-only cache, and the f64 FIR pass
// of fir_f64.cu (launch_fir_pass_f64) for the two filtfilt passes.
#include "common.cuh"
#include "ext_modes.cuh"

#include <algorithm>
#include <numeric>

namespace scir_b200 {

namespace {

// ---- upfirdn, f64, any rate ----------------------------------------------------------------------------------------
// y[m] = sum_i h[t + i*up] * xe[q - i],  t = (m*down) % up,  q = (m*down) / up   (pyx:421-481), oldest sample first.
// A CTA owns kUpTile consecutive outputs of one row: it stages the input span they touch (extension applied while
// staging) and walks the phase-transposed taps ht[t][i] = h[t + i*up] from global memory (L1-resident: up * hpp doubles).
constexpr int kUpThreads = 256;
constexpr int kUpPer = 4;                                   // outputs per thread
constexpr int kUpTile = kUpThreads * kUpPer;

struct Up64Params {
    const double* x;
    double* y;
    const double* ht;            // [up][hpp]
    long long ld_x, ld_y, n_in;
    long long m_begin, m_end;
    long long tiles_per_row;
    long long up, down;
    int hpp;                     // taps per phase
    int span_cap;                // doubles reserved for the input span
    ExtSpec64 ext;
};

__global__ void __launch_bounds__(kUpThreads) upfirdn_f64_kernel(const __grid_constant__ Up64Params q)
{
    extern __shared__ double xs64[];
    const long long row = blockIdx.x / q.tiles_per_row;
    const long long tile = blockIdx.x - row * q.tiles_per_row;
    const long long m0 = q.m_begin + tile * kUpTile;
    const long long m_last = min(m0 + kUpTile, q.m_end) - 1;
    const long long q_first = (m0 * q.down) / q.up - (q.hpp - 1);
    const int span = static_cast<int>((m_last * q.down) / q.up - q_first + 1);
    const double* __restrict__ xr = q.x + row * q.ld_x;
    const bool zpad = (q.ext.mode == SCIR_B200_EXT_CONSTANT && q.ext.cval == 0.0);
    for (int s = threadIdx.x; s < span; s += kUpThreads) {
        const long long xi = q_first + s;
        xs64[s] = (xi >= 0 && xi < q.n_in) ? xr[xi] : (zpad ? 0.0 : upfirdn_sample(xr, xi, q.n_in, q.ext));
    }
    __syncthreads();
    double* __restrict__ yr = q.y + row * q.ld_y;
#pragma unroll
    for (int j = 0; j < kUpPer; ++j) {
        const long long m = m0 + threadIdx.x + static_cast<long long>(j) * kUpThreads;     // coalesced stores
        if (m >= q.m_end) continue;
        const long long md = m * q.down;
        const int t = static_cast<int>(md % q.up);
        const int qrel = static_cast<int>(md / q.up - q_first);
        const double* __restrict__ trow = q.ht + static_cast<long long>(t) * q.hpp;
        double acc = 0.0;
        for (int i = q.hpp - 1; i >= 0; --i) acc = fma(trow[i], xs64[qrel - i], acc);     // oldest first (pyx:451-453)
        yr[m - q.m_begin] = acc;
    }
}

// ---- per-row statistics in f64 (resample_poly padtype = mean / minimum / maximum / median) ----------------------------
template <int STAT>
__global__ void __launch_bounds__(512) row_stat_f64_kernel(const double* __restrict__ x, long long ld_x, long long n, double* __restrict__ out)
{
    const double* xr = x + static_cast<long long>(blockIdx.x) * ld_x;
    double v = (STAT == 0) ? 0.0 : xr[0];
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double t = xr[i];
        v = (STAT == 0) ? (v + t) : (STAT == 1 ? fmin(v, t) : fmax(v, t));
    }
    __shared__ double sh[512];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = 256; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double a = sh[threadIdx.x], b = sh[threadIdx.x + s];
            sh[threadIdx.x] = (STAT == 0) ? (a + b) : (STAT == 1 ? fmin(a, b) : fmax(a, b));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (STAT == 0) ? sh[0] / static_cast<double>(n) : sh[0];
}

// exact k-th order statistic by 8 passes of an 8-bit radix select over order-preserving 64-bit keys
__device__ __forceinline__ unsigned long long f64_sort_key(double v)
{
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_sort_key(unsigned long long k)
{
    return __longlong_as_double(static_cast<long long>((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

__device__ double row_select_f64(const double* __restrict__ xr, long long n, long long kth, unsigned* hist, unsigned long long* state)
{
    unsigned long long prefix = 0, mask = 0;
    long long k = kth;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = f64_sort_key(xr[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255ull], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long acc = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (acc + hist[b] > k) break;
                acc += hist[b];
            }
            state[0] = static_cast<unsigned long long>(b);
            state[1] = static_cast<unsigned long long>(k - acc);
        }
        __syncthreads();
        prefix |= state[0] << shift;
        mask |= 255ull << shift;
        k = static_cast<long long>(state[1]);
        __syncthreads();
    }
    return f64_from_sort_key(prefix);
}

__global__ void __launch_bounds__(1024) row_median_f64_kernel(const double* __restrict__ x, long long ld_x, long long n, double* __restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long state[2];
    const double* xr = x + static_cast<long long>(blockIdx.x) * ld_x;
    const double hi = row_select_f64(xr, n, n / 2, hist, state);
    double med = hi;
    if ((n & 1) == 0) {
        const double lo = row_select_f64(xr, n, n / 2 - 1, hist, state);
        med = __dmul_rn(__dadd_rn(lo, hi), 0.5);            // numpy: mean of the two middle values
    }
    if (threadIdx.x == 0) out[blockIdx.x] = med;
}

// y[row, i] = x[row, i] + sign * bg[row]   (y may alias x)
__global__ void __launch_bounds__(256) row_offset_f64_kernel(const double* x, long long ld_x, const double* __restrict__ bg, double sign,
                                                             double* y, long long ld_y, long long n)
{
    const long long row = blockIdx.y;
    const double b = sign * bg[row];
    const double* xr = x + row * ld_x;
    double* yr = y + row * ld_y;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        yr[i] = __dadd_rn(xr[i], b);
}

int check_matrix64(const void* p, int64_t ld, int64_t batch, int64_t n, const char* name)
{
    if (batch < 0 || n < 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s: negative shape", name);
    if (batch > 0 && n > 0 && p == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s is NULL", name);
    if (batch > 1 && ld < n) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s: ld (%lld) < row length (%lld)", name, (long long)ld, (long long)n);
    return SCIR_B200_OK;
}

int check_no_alias64(const double* x, int64_t ld_x, int64_t nx, const double* y, int64_t ld_y, int64_t ny, int64_t batch)
{
    if (batch <= 0 || nx <= 0 || ny <= 0 || x == nullptr || y == nullptr) return SCIR_B200_OK;
    const double* xe = x + (batch - 1) * ld_x + nx;
    const double* ye = y + (batch - 1) * ld_y + ny;
    if (x < ye && y < xe) return set_error(SCIR_B200_ERR_INVALID_ARG, "x and y overlap: the FIR routes cannot run in place");
    return SCIR_B200_OK;
}

int launch_upfirdn_f64(scir_b200_ctx* ctx, const double* h, int64_t len_h, int64_t up, int64_t down, const double* d_x, int64_t ld_x,
                       int64_t batch, int64_t n_in, double* d_y, int64_t ld_y, int64_t m_begin, int64_t m_count, int ext_mode, double cval)
{
    if (ext_mode < SCIR_B200_EXT_CONSTANT || ext_mode > SCIR_B200_EXT_LINE)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown extension mode %d", ext_mode);
    const bool needs2 = ext_mode == SCIR_B200_EXT_REFLECT || ext_mode == SCIR_B200_EXT_SMOOTH || ext_mode == SCIR_B200_EXT_LINE ||
                        ext_mode == SCIR_B200_EXT_ANTIREFLECT;
    if (needs2 && n_in < 2) return set_error(SCIR_B200_ERR_SHAPE, "extension mode %d needs at least two samples per row", ext_mode);
    if (batch == 0 || m_count == 0) return SCIR_B200_OK;
    const int64_t hpp = (len_h + up - 1) / up;
    const int64_t span_cap = static_cast<int64_t>(kUpTile) * down / up + hpp + 4;
    const size_t smem = static_cast<size_t>(span_cap) * sizeof(double);
    if (hpp > (1 << 20) || smem > static_cast<size_t>(ctx->max_smem_optin))
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "f64 upfirdn: rate %lld/%lld with %lld taps needs %zu B of shared memory",
                         (long long)up, (long long)down, (long long)len_h, smem);
    const int64_t tiles = (m_count + kUpTile - 1) / kUpTile;
    if (tiles * batch > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "grid too large");
    SCIR_TRY(ctx_bind(ctx));
    std::vector<double> ht(static_cast<size_t>(up * hpp), 0.0);
    for (int64_t t = 0; t < up; ++t)
        for (int64_t i = 0; i < hpp; ++i)
            if (t + i * up < len_h) ht[static_cast<size_t>(t * hpp + i)] = h[t + i * up];
    double* d_ht = nullptr;
    SCIR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_ht), ht.size() * sizeof(double), ctx->stream), "cudaMallocAsync(taps)");
    // pageable source: cudaMemcpyAsync returns once the bytes sit in the driver's staging buffer, so `ht` may die with this frame
    SCIR_CUDA(cudaMemcpyAsync(d_ht, ht.data(), ht.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(taps)");
    Up64Params q{};
    q.x = d_x; q.y = d_y; q.ht = d_ht; q.ld_x = ld_x; q.ld_y = ld_y; q.n_in = n_in;
    q.m_begin = m_begin; q.m_end = m_begin + m_count; q.tiles_per_row = tiles; q.up = up; q.down = down;
    q.hpp = static_cast<int>(hpp); q.span_cap = static_cast<int>(span_cap); q.ext = ExtSpec64{ext_mode, cval};
    static thread_local size_t configured[16] = {};
    const int d = ctx->device & 15;
    if (smem > 48 * 1024 && configured[d] < smem) {
        SCIR_CUDA(cudaFuncSetAttribute(upfirdn_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "cudaFuncSetAttribute(upfirdn_f64_kernel)");
        configured[d] = smem;
    }
    upfirdn_f64_kernel<<<static_cast<unsigned>(tiles * batch), kUpThreads, smem, ctx->stream>>>(q);
    SCIR_CUDA(cudaGetLastError(), "upfirdn_f64_kernel launch");
    ctx->launches++;
    SCIR_CUDA(cudaFreeAsync(d_ht, ctx->stream), "cudaFreeAsync(taps)");
    return SCIR_B200_OK;
}

int resample_plan64(int64_t n_in, int64_t len_h, int64_t up, int64_t down, scir_b200_resample_plan* p)
{
    return scir_b200_resample_poly_plan(n_in, len_h, up, down, p);
}

int resample_device_f64(scir_b200_ctx* ctx, const double* window, int64_t len_h, int64_t up, int64_t down, const double* d_x,
                        int64_t ld_x, int64_t batch, int64_t n_in, double* d_y, int64_t ld_y, int ext_mode, double cval)
{
    scir_b200_resample_plan pl;
    SCIR_TRY(resample_plan64(n_in, len_h, up, down, &pl));
    if (batch == 0 || n_in == 0) return SCIR_B200_OK;
    if (pl.up == 1 && pl.down == 1) {
        SCIR_TRY(ctx_bind(ctx));
        SCIR_CUDA(cudaMemcpy2DAsync(d_y, ld_y * sizeof(double), d_x, ld_x * sizeof(double), n_in * sizeof(double), batch,
                                    cudaMemcpyDeviceToDevice, ctx->stream),
                  "cudaMemcpy2DAsync(resample copy)");
        return SCIR_B200_OK;
    }
    // h = window * up (:3909), zero-padded front and back (:3919-3920)
    std::vector<double> h(static_cast<size_t>(pl.len_h_padded), 0.0);
    for (int64_t i = 0; i < len_h; ++i) h[static_cast<size_t>(pl.n_pre_pad + i)] = window[i] * static_cast<double>(pl.up);
    return launch_upfirdn_f64(ctx, h.data(), pl.len_h_padded, pl.up, pl.down, d_x, ld_x, batch, n_in, d_y, ld_y, pl.n_pre_remove,
                              pl.n_out, ext_mode, cval);
}

int filtfilt_device_f64(scir_b200_ctx* ctx, const double* b, int64_t k, int pad_mode, int64_t padlen, const double* d_x, int64_t ld_x,
                        double* d_y, int64_t ld_y, int64_t batch, int64_t n)
{
    int64_t edge = 0;
    int ext = EXT_NONE;
    int bound = BOUND_HOLD;
    switch (pad_mode) {
        case SCIR_B200_PAD_ZERO_STATE: bound = BOUND_ZERO; break;
        case SCIR_B200_PAD_SCIPY_NONE: break;
        case SCIR_B200_PAD_ODD: ext = EXT_ODD; edge = padlen < 0 ? 3 * k : padlen; break;
        case SCIR_B200_PAD_EVEN: ext = EXT_EVEN; edge = padlen < 0 ? 3 * k : padlen; break;
        case SCIR_B200_PAD_CONSTANT: ext = EXT_CONST; edge = padlen < 0 ? 3 * k : padlen; break;
        default: return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown pad_mode %d", pad_mode);
    }
    if (pad_mode != SCIR_B200_PAD_ZERO_STATE && n <= edge)
        return set_error(SCIR_B200_ERR_SHAPE, "The length of the input vector x must be greater than padlen, which is %lld.",
                         (long long)edge);                       // _signaltools.py:4809
    if (batch == 0 || n == 0) return SCIR_B200_OK;
    if (edge == 0) ext = EXT_NONE;
    // Two passes exactly as SciPy structures them (forward over the extended signal with steady-state initial
    // conditions, then backward): f64 callers want SciPy's values to ~1e-15, so no single-pass fusion here.
    const int64_t n_v = n + 2 * edge;
    SCIR_TRY(ctx_scratch(ctx, ctx->scratch, static_cast<size_t>(batch) * n_v * sizeof(double)));
    double* y1 = static_cast<double*>(ctx->scratch.ptr);
    FirPass64 f{};
    f.x = d_x; f.ld_x = ld_x; f.y = y1; f.ld_y = n_v; f.batch = batch;
    f.n_x = n; f.n_v = n_v; f.in_off = -edge; f.out_off = 0;
    f.out_begin = 0; f.out_end = n_v; f.ext_mode = ext; f.bound = bound; f.dir = +1;
    SCIR_TRY(launch_fir_pass_f64(ctx, f, b, k));
    FirPass64 r{};
    r.x = y1; r.ld_x = n_v; r.y = d_y; r.ld_y = ld_y; r.batch = batch;
    r.n_x = n_v; r.n_v = n_v; r.in_off = 0; r.out_off = -edge;
    r.out_begin = edge; r.out_end = edge + n; r.ext_mode = EXT_NONE; r.bound = bound; r.dir = -1;
    return launch_fir_pass_f64(ctx, r, b, k);
}

}  // namespace

}  // namespace scir_b200

using namespace scir_b200;

extern "C" {

int scir_b200_upfirdn_mode_f64(scir_b200_ctx* ctx, const double* h, int64_t len_h, int64_t up, int64_t down, int mode, double cval,
                               const double* d_x, int64_t ld_x, int64_t batch, int64_t n_in, double* d_y, int64_t ld_y,
                               int64_t m_begin, int64_t m_count)
{
    SCIR_ENTER(ctx);
    if (!h || len_h < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "h must hold at least one tap");
    if (up < 1 || down < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "up and down must be >= 1");   // _upfirdn.py:98
    if (n_in < 1 && batch > 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "upfirdn needs n_in >= 1");
    SCIR_TRY(check_matrix64(d_x, ld_x, batch, n_in, "x"));
    if (m_begin < 0 || m_count < 0 || (batch > 0 && m_begin + m_count > upfirdn_out_len(len_h, n_in, up, down)))
        return set_error(SCIR_B200_ERR_INVALID_ARG, "output window [%lld, %lld) outside the upfirdn result", (long long)m_begin,
                         (long long)(m_begin + m_count));
    SCIR_TRY(check_matrix64(d_y, ld_y, batch, m_count, "y"));
    SCIR_TRY(check_no_alias64(d_x, ld_x, n_in, d_y, ld_y, m_count, batch));
    return launch_upfirdn_f64(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count, mode, cval);
}

int scir_b200_resample_poly_pad_f64(scir_b200_ctx* ctx, const double* window, int64_t len_h, int64_t up, int64_t down, int padtype,
                                    double cval, const double* d_x, int64_t ld_x, int64_t batch, int64_t n_in, double* d_y,
                                    int64_t ld_y)
{
    SCIR_ENTER(ctx);
    scir_b200_resample_plan pl;
    if (!window) return set_error(SCIR_B200_ERR_INVALID_ARG, "window is NULL");
    SCIR_TRY(scir_b200_resample_poly_plan(n_in, len_h, up, down, &pl));
    SCIR_TRY(check_matrix64(d_x, ld_x, batch, n_in, "x"));
    const bool copy = (pl.up == 1 && pl.down == 1);
    const int64_t n_out = copy ? n_in : pl.n_out;
    SCIR_TRY(check_matrix64(d_y, ld_y, batch, n_out, "y"));
    SCIR_TRY(check_no_alias64(d_x, ld_x, n_in, d_y, ld_y, n_out, batch));
    if (padtype >= SCIR_B200_EXT_CONSTANT && padtype <= SCIR_B200_EXT_LINE)
        return resample_device_f64(ctx, window, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, padtype, cval);
    if (padtype != SCIR_B200_PAD_STAT_MEAN && padtype != SCIR_B200_PAD_STAT_MINIMUM && padtype != SCIR_B200_PAD_STAT_MAXIMUM &&
        padtype != SCIR_B200_PAD_STAT_MEDIAN)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown padtype %d", padtype);
    if (batch == 0 || n_in == 0 || copy)          // SciPy returns x.copy() before looking at padtype (:3885-3886)
        return resample_device_f64(ctx, window, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, SCIR_B200_EXT_CONSTANT, 0.0);
    if (batch > 0x7fffffffLL) return set_error(SCIR_B200_ERR_UNSUPPORTED, "too many rows");
    // background statistic per row, x - bg into scratch, zero-padded upfirdn, + bg (:3927-3957)
    SCIR_TRY(ctx_scratch(ctx, ctx->row_bg, static_cast<size_t>(batch) * sizeof(double)));
    SCIR_TRY(ctx_scratch(ctx, ctx->scratch, static_cast<size_t>(batch) * n_in * sizeof(double)));
    double* bg = static_cast<double*>(ctx->row_bg.ptr);
    double* xc = static_cast<double*>(ctx->scratch.ptr);
    const unsigned g = static_cast<unsigned>(batch);
    if (padtype == SCIR_B200_PAD_STAT_MEDIAN) row_median_f64_kernel<<<g, 1024, 0, ctx->stream>>>(d_x, ld_x, n_in, bg);
    else if (padtype == SCIR_B200_PAD_STAT_MEAN) row_stat_f64_kernel<0><<<g, 512, 0, ctx->stream>>>(d_x, ld_x, n_in, bg);
    else if (padtype == SCIR_B200_PAD_STAT_MINIMUM) row_stat_f64_kernel<1><<<g, 512, 0, ctx->stream>>>(d_x, ld_x, n_in, bg);
    else row_stat_f64_kernel<2><<<g, 512, 0, ctx->stream>>>(d_x, ld_x, n_in, bg);
    SCIR_CUDA(cudaGetLastError(), "row statistic (f64) launch");
    ctx->launches++;
    auto offset = [&](const double* src, int64_t ld_s, double sign, double* dst, int64_t ld_d, int64_t n) -> int {
        for (int64_t r0 = 0; r0 < batch; r0 += 65535) {
            const int64_t nr = std::min<int64_t>(65535, batch - r0);
            const unsigned gx = static_cast<unsigned>(std::min<long long>((n + 255) / 256, 64));
            row_offset_f64_kernel<<<dim3(gx, static_cast<unsigned>(nr)), 256, 0, ctx->stream>>>(src + r0 * ld_s, ld_s, bg + r0, sign,
                                                                                                  dst + r0 * ld_d, ld_d, n);
            SCIR_CUDA(cudaGetLastError(), "row offset (f64) launch");
            ctx->launches++;
        }
        return SCIR_B200_OK;
    };
    SCIR_TRY(offset(d_x, ld_x, -1.0, xc, n_in, n_in));
    SCIR_TRY(resample_device_f64(ctx, window, len_h, up, down, xc, n_in, batch, n_in, d_y, ld_y, SCIR_B200_EXT_CONSTANT, 0.0));
    return offset(d_y, ld_y, +1.0, d_y, ld_y, n_out);
}

int scir_b200_filtfilt_fir_f64(scir_b200_ctx* ctx, const double* b, int64_t k, int pad_mode, int64_t padlen, const double* d_x,
                               int64_t ld_x, double* d_y, int64_t ld_y, int64_t batch, int64_t n)
{
    SCIR_ENTER(ctx);
    if (b == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "taps is NULL");
    if (k < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "need at least one tap (k=%lld)", (long long)k);
    SCIR_TRY(check_matrix64(d_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix64(d_y, ld_y, batch, n, "y"));
    SCIR_TRY(check_no_alias64(d_x, ld_x, n, d_y, ld_y, n, batch));
    return filtfilt_device_f64(ctx, b, k, pad_mode, padlen, d_x, ld_x, d_y, ld_y, batch, n);
}

}  // extern "C"
