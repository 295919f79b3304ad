// api.cu -- C-ABI entry points (include/scir_b200.h): errors, context, memory, and the host-side
// dispatch that maps scir-signal's FIR routes onto the kernels.  This is the C++ host layer that
// stands where the reference's `mod cuda` host wrappers stood (crates/scir-gpu/src/lib.rs:840-1113).
#include "common.cuh"
#include "copy_pool.hpp"

#include <cmath>

#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <numeric>
#include <thread>

namespace scir_b200 {

// ---- errors -------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void clear_error() { g_err[0] = '\0'; }

int cuda_error(cudaError_t e, const char* what)
{
    int code = SCIR_B200_ERR_LAUNCH;
    if (e == cudaErrorMemoryAllocation) code = SCIR_B200_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice ||
        e == cudaErrorInitializationError || e == cudaErrorSystemDriverMismatch ||
        e == cudaErrorSystemNotReady || e == cudaErrorNotSupported)
        code = SCIR_B200_ERR_NO_DEVICE;
    // a sticky error must not poison unrelated later calls' diagnostics
    cudaGetLastError();
    return set_error(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

int check_ctx(const scir_b200_ctx* ctx)
{
    if (ctx == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "ctx is NULL");
    return SCIR_B200_OK;
}

int ctx_bind(const scir_b200_ctx* ctx)
{
    SCIR_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    return SCIR_B200_OK;
}

// ---- DeviceScope: the caller's current context survives every ABI call ---------------------------
namespace {
typedef int (*CtxGetCurrentFn)(void**);
typedef int (*CtxSetCurrentFn)(void*);
struct DriverCtxApi {
    CtxGetCurrentFn get = nullptr;
    CtxSetCurrentFn set = nullptr;
    DriverCtxApi()
    {
        void *g = nullptr, *s = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &g, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuCtxSetCurrent", &s, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess) {
            get = reinterpret_cast<CtxGetCurrentFn>(g);
            set = reinterpret_cast<CtxSetCurrentFn>(s);
        } else {
            cudaGetLastError();
        }
    }
};
const DriverCtxApi& driver_ctx_api()
{
    static const DriverCtxApi api;
    return api;
}
}  // namespace

DeviceScope::DeviceScope(int device)
{
    const DriverCtxApi& api = driver_ctx_api();
    if (api.get && api.get(&prev) == 0) restore = true;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) rc = cuda_error(e, "cudaSetDevice");
}

DeviceScope::~DeviceScope()
{
    if (!restore) return;
    const DriverCtxApi& api = driver_ctx_api();
    void* now = nullptr;
    if (api.get(&now) == 0 && now != prev) api.set(prev);
}

int ctx_scratch(scir_b200_ctx* ctx, DeviceBuffer& buf, size_t bytes)
{
    if (buf.bytes >= bytes && buf.ptr != nullptr) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    if (buf.ptr) {
        // the old buffer may still be in use by queued work on any of the ctx's streams
        SCIR_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
        if (ctx->s_h2d) SCIR_CUDA(cudaStreamSynchronize(ctx->s_h2d), "cudaStreamSynchronize");
        if (ctx->s_d2h) SCIR_CUDA(cudaStreamSynchronize(ctx->s_d2h), "cudaStreamSynchronize");
        SCIR_CUDA(cudaFree(buf.ptr), "cudaFree(scratch)");
        buf.ptr = nullptr;
        buf.bytes = 0;
    }
    SCIR_CUDA(cudaMalloc(&buf.ptr, bytes), "cudaMalloc(scratch)");
    buf.bytes = bytes;
    return SCIR_B200_OK;
}

int64_t upfirdn_out_len(int64_t len_h, int64_t in_len, int64_t up, int64_t down)
{
    // scipy/signal/_upfirdn_apply.pyx:59-67 (floor division; operands are positive here)
    return (((in_len - 1) * up + len_h) - 1) / down + 1;
}

static int check_matrix(const void* p, int64_t ld, int64_t batch, int64_t n, const char* name)
{
    if (batch < 0 || n < 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s: negative shape", name);
    if (batch > 0 && n > 0 && p == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s is NULL", name);
    if (batch > 1 && ld < n) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s: ld (%lld) < row length (%lld)",
                                              name, (long long)ld, (long long)n);
    return SCIR_B200_OK;
}

// The kernels read a tile plus its halo while other CTAs write their outputs: filtering in place is undefined.
// (The reference always returns a fresh array, lib.rs:1044.)  Elementwise ops are the exception and say so.
template <typename T>
static int check_no_alias(const T* x, int64_t ld_x, int64_t nx, const T* y, int64_t ld_y, int64_t ny, int64_t batch)
{
    if (batch <= 0 || nx <= 0 || ny <= 0 || x == nullptr || y == nullptr) return SCIR_B200_OK;
    const T* xe = x + (batch - 1) * ld_x + nx;
    const T* ye = y + (batch - 1) * ld_y + ny;
    if (x < ye && y < xe)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "x and y overlap: the FIR routes cannot run in place");
    return SCIR_B200_OK;
}

static int check_taps(const float* taps, int64_t k)
{
    if (taps == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "taps is NULL");
    if (k < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "need at least one tap (k=%lld)", (long long)k);
    if (k > SCIR_B200_MAX_TAPS)
        return set_error(SCIR_B200_ERR_UNSUPPORTED, "k=%lld exceeds SCIR_B200_MAX_TAPS=%d", (long long)k,
                         SCIR_B200_MAX_TAPS);
    return SCIR_B200_OK;
}

// Kernel choice for one FIR pass, from measurements on B200 (profiles/README.md):
//   * the FP32 direct family is FFMA-issue bound, time ~ K: at K = 63 it needs 2.40 ms for config 2 where
//     the bytes alone take 1.31 ms;
//   * the tcgen05 block-Toeplitz contraction (block-scaled FP16x3, error <= 3 * 2^-22 per product) runs the
//     same launch in 1.5 ms and stays HBM-bound up to K ~ 130 -- but it is one persistent CTA per SM with a
//     ~20 us floor (TMEM allocation, Hankel tap arrays, a three-deep pipeline to fill, the fix-up launch),
//     and it works in whole 128 x 128 output tiles.
// So both are costed with a small model (fitted to the sweep in profiles/README.md) and the cheaper one runs;
// K >= toeplitz_min_k (1024) always takes the tensor path, where the direct kernel is 7x slower at any size
// that matters.  The direct kernel's own streaming floor (~5.2 TB/s) means large launches go to the tensor
// path even for short filters: there it is simply the better streaming kernel.
// the tensor kernel's cost: max(20 us, 16 us + rounds * max(3.6 us [HBM share of one tile per SM], 58 ns per MMA))
static double toeplitz_seconds(const scir_b200_ctx* ctx, int64_t k, int64_t tiles)
{
    const int64_t pmax = (k - 1 + 127) / 128;
    int64_t ksteps = 0;
    for (int64_t pb = 0; pb <= pmax; ++pb) ksteps += 8 - (std::max<int64_t>(0, 128 * pb - (k - 1)) >> 4);
    const double t_round = std::max(3.6e-6, static_cast<double>(3 * ksteps) * 58e-9);
    const double rounds = std::ceil(static_cast<double>(tiles) / static_cast<double>(ctx->sm_count));
    return std::max(20e-6, 16e-6 + rounds * t_round);
}

static bool prefer_toeplitz(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k, int64_t tiles, bool aligned)
{
    // fitted to tools/sweep_dispatch.py on B200 (45 shapes, K = 48 .. 511, 64 .. 16384 tiles; profiles/README.md):
    //   direct:   14 us + per 16384 outputs max(25 ns [its streaming floor, ~5.2 TB/s], (2K + 29) flop / 72 TFLOP/s)
    const double units = static_cast<double>(pass.out_end - pass.out_begin) * static_cast<double>(pass.batch) / 16384.0;
    const double t_direct = 14e-6 + units * std::max(25e-9, 16384.0 * (2.0 * static_cast<double>(k) + 29.0) / 72e12);
    const double t_toep = toeplitz_seconds(ctx, k, tiles);
    // rows that are not 16-byte aligned (odd views / pitches) lose the bulk-copy staging in both families: measured
    // (tools/time_unaligned.py, 256 x 2^18) x1.7 on the tensor kernel, x1.0-1.4 on the direct kernel
    return aligned ? (t_toep < t_direct) : (1.7 * t_toep < 1.2 * t_direct);
}

int launch_fir(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k)
{
    const int64_t mode = ctx->opt.long_tap_path;
    // Long filters: block-FFT convolution (fir_os.cu) does O(log N) work per output where direct form does O(K).
    // Measured cross-over against the tensor kernel: profiles/README.md.  A launch must fill the machine with block
    // pairs (2 (N - K + 1) outputs each) for the FFT path to pay.
    double t_os = 0.0;
    if ((mode == 3 || (mode == 0 && ctx->opt.variant == 0 && k >= ctx->opt.os_min_k)) && fir_os_supported(ctx, pass, k, &t_os)) {
        if (mode == 3) return launch_fir_os(ctx, pass, c, k);
        int64_t tiles = 0;
        bool aligned = false;
        // against the tensor kernel's model (the FFT path reads rows with scalar loads: alignment does not matter to it)
        if (!toeplitz_supported(ctx, pass, k, &tiles, &aligned) || t_os < (aligned ? 1.0 : 1.7) * toeplitz_seconds(ctx, k, tiles))
            return launch_fir_os(ctx, pass, c, k);
    }
    if (mode != 1 && mode != 3 && ctx->opt.variant == 0) {
        int64_t tiles = 0;
        bool aligned = false;
        if (toeplitz_supported(ctx, pass, k, &tiles, &aligned)) {
            const bool want = (mode == 2) || k >= ctx->opt.toeplitz_min_k ||
                              (k >= ctx->opt.toeplitz_min_k_full && prefer_toeplitz(ctx, pass, k, tiles, aligned));
            if (want) return launch_fir_toeplitz(ctx, pass, c, k);
        }
    }
    return launch_fir_pass(ctx, pass, c, k);
}

// One causal zero-state FIR over (batch, n): the reference hot path.
static int fir_causal(scir_b200_ctx* ctx, const float* c, int64_t k, const float* d_x, int64_t ld_x,
                      float* d_y, int64_t ld_y, int64_t batch, int64_t n)
{
    FirPass p{};
    p.x = d_x; p.y = d_y; p.ld_x = ld_x; p.ld_y = ld_y; p.batch = batch;
    p.n_x = n; p.n_v = n; p.in_off = 0; p.out_off = 0; p.out_begin = 0; p.out_end = n;
    p.ext_mode = EXT_NONE; p.bound = BOUND_ZERO; p.dir = +1;
    return launch_fir(ctx, p, c, k);
}

static int filtfilt_device(scir_b200_ctx* ctx, const float* b, int64_t k, int pad_mode, int64_t padlen,
                           const float* d_x, int64_t ld_x, float* d_y, int64_t ld_y, int64_t batch,
                           int64_t n)
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
        return set_error(SCIR_B200_ERR_SHAPE,
                         "The length of the input vector x must be greater than padlen, which is %lld.",
                         (long long)edge);                       // _signaltools.py:4809
    if (batch == 0 || n == 0) return SCIR_B200_OK;
    if (edge == 0) ext = EXT_NONE;

    const int64_t n_v = n + 2 * edge;

    // ---- one pass instead of two ----------------------------------------------------------------------------
    // With an extension of at least k-1 samples (SciPy's default is 3k) no kept output ever sees the held boundary
    // of either pass: y[i] = sum_d b[d] * y1[i+d], y1[j] = sum_e b[e] * ext[j-e] with every index inside the
    // extended signal, i.e. ONE zero-phase filter hc = b (*) flip(b) of 2k-1 taps over ext.  Same values to
    // rounding (hc is formed in f64), half the HBM traffic, no intermediate matrix, and on the tensor path
    // fewer MMAs than two k-tap passes (k = 255: 120 vs 144 per tile).  filtfilt_fused=0 keeps the two-pass form.
    if (ext != EXT_NONE && edge >= k - 1 && 2 * k - 1 <= SCIR_B200_MAX_TAPS && ctx->opt.filtfilt_fused != 0 &&
        (2 * k - 1 <= 4097 || k > 4097)) {
        const int64_t kc = 2 * k - 1;
        std::vector<float> hc(static_cast<size_t>(kc));
        for (int64_t m = 0; m < kc; ++m) {                        // hc[m] = sum_j b[j] * b[j + (k-1) - m]
            double acc = 0.0;
            for (int64_t j = 0; j < k; ++j) {
                const int64_t j2 = j + (k - 1) - m;
                if (j2 >= 0 && j2 < k) acc += static_cast<double>(b[j]) * static_cast<double>(b[j2]);
            }
            hc[static_cast<size_t>(m)] = static_cast<float>(acc);
        }
        FirPass f{};                                              // causal 2k-1 taps, output delayed by k-1: centred
        f.x = d_x; f.ld_x = ld_x; f.y = d_y; f.ld_y = ld_y; f.batch = batch;
        f.n_x = n; f.n_v = n_v; f.in_off = -edge; f.out_off = -(edge + (k - 1));
        f.out_begin = edge + (k - 1); f.out_end = edge + (k - 1) + n;
        f.ext_mode = ext; f.bound = bound; f.dir = +1;
        ctx->filtfilt_fused_calls++;
        return launch_fir(ctx, f, hc.data(), kc);
    }

    const int64_t padlead = (4 - edge % 4) % 4;                   // keeps both passes 16-B aligned
    const int64_t ld1 = (n_v + padlead + 3) / 4 * 4;
    SCIR_TRY(ctx_scratch(ctx, ctx->scratch, static_cast<size_t>(batch) * ld1 * sizeof(float)));
    float* y1 = static_cast<float*>(ctx->scratch.ptr);

    FirPass f{};                                                  // forward over the extended signal
    f.x = d_x; f.ld_x = ld_x; f.y = y1; f.ld_y = ld1; f.batch = batch;
    f.n_x = n; f.n_v = n_v; f.in_off = -edge; f.out_off = padlead;
    f.out_begin = 0; f.out_end = n_v; f.ext_mode = ext; f.bound = bound; f.dir = +1;
    SCIR_TRY(launch_fir(ctx, f, b, k));

    FirPass r{};                                                  // backward, keep [edge, edge+n)
    r.x = y1; r.ld_x = ld1; r.y = d_y; r.ld_y = ld_y; r.batch = batch;
    r.n_x = n_v; r.n_v = n_v; r.in_off = padlead; r.out_off = -edge;
    r.out_begin = edge; r.out_end = edge + n; r.ext_mode = EXT_NONE; r.bound = bound; r.dir = -1;
    return launch_fir(ctx, r, b, k);
}

static int resample_plan(int64_t n_in, int64_t len_h, int64_t up, int64_t down, scir_b200_resample_plan* p)
{
    // scipy/signal/_signaltools.py:3882-3918, int64 throughout
    const int64_t g = std::gcd(up, down);
    up /= g;
    down /= g;
    p->up = up;
    p->down = down;
    const int64_t prod = n_in * up;
    p->n_out = prod / down + (prod % down != 0 ? 1 : 0);
    p->half_len = (len_h - 1) / 2;
    p->n_pre_pad = down - p->half_len % down;
    p->n_post_pad = 0;
    p->n_pre_remove = (p->half_len + p->n_pre_pad) / down;
    while (upfirdn_out_len(len_h + p->n_pre_pad + p->n_post_pad, n_in, up, down) < p->n_out + p->n_pre_remove)
        p->n_post_pad += 1;
    p->len_h_padded = len_h + p->n_pre_pad + p->n_post_pad;
    p->upfirdn_len = upfirdn_out_len(p->len_h_padded, n_in, up, down);
    return SCIR_B200_OK;
}

static int resample_device(scir_b200_ctx* ctx, const float* window, int64_t len_h, int64_t up, int64_t down,
                           const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y,
                           int64_t ld_y, int ext_mode = 0, float cval = 0.f)
{
    scir_b200_resample_plan pl;
    resample_plan(n_in, len_h, up, down, &pl);
    if (batch == 0 || n_in == 0) return SCIR_B200_OK;
    if (pl.up == 1 && pl.down == 1) {
        SCIR_TRY(ctx_bind(ctx));
        SCIR_CUDA(cudaMemcpy2DAsync(d_y, ld_y * sizeof(float), d_x, ld_x * sizeof(float), n_in * sizeof(float),
                                    batch, cudaMemcpyDeviceToDevice, ctx->stream),
                  "cudaMemcpy2DAsync(resample copy)");
        return SCIR_B200_OK;
    }
    // h = window * up in f32 (:3909), zero-padded front and back (:3919-3920)
    std::vector<float> h(static_cast<size_t>(pl.len_h_padded), 0.f);
    for (int64_t i = 0; i < len_h; ++i) h[static_cast<size_t>(pl.n_pre_pad + i)] = window[i] * static_cast<float>(pl.up);
    return launch_upfirdn(ctx, h.data(), pl.len_h_padded, pl.up, pl.down, d_x, ld_x, batch, n_in, d_y, ld_y,
                          pl.n_pre_remove, pl.n_out, ext_mode, cval);
}

// ---- *_host streaming: rows flow through a ring of row blocks, H2D / kernels / D2H on three streams --------
// Pinned caller memory (scir_b200_host_alloc / scir_b200_host_register) is DMA'd directly.  Pageable caller memory
// -- what a Rust Vec / ndarray::Array2 / numpy array is -- would make cudaMemcpyAsync stage through the driver's
// own bounce buffer on the calling thread, which serialises the three streams; instead the ring gets pinned
// twins, filled and drained by the ctx's copy threads while the DMA engines and the kernels run (host_stage=1),
// or the caller's spans are registered for the duration of the call (host_stage=2).
static bool is_pinned_span(const void* p, size_t bytes)
{
    if (!p || bytes == 0) return true;
    auto pinned = [](const void* q) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
    };
    return pinned(p) && pinned(static_cast<const char*>(p) + bytes - 1);
}

static int pinned_scratch(scir_b200_ctx* ctx, DeviceBuffer& buf, size_t bytes, unsigned flags)
{
    if (buf.bytes >= bytes && buf.ptr != nullptr) return SCIR_B200_OK;
    if (buf.ptr) {
        SCIR_CUDA(cudaStreamSynchronize(ctx->s_h2d), "cudaStreamSynchronize");
        SCIR_CUDA(cudaStreamSynchronize(ctx->s_d2h), "cudaStreamSynchronize");
        SCIR_CUDA(cudaFreeHost(buf.ptr), "cudaFreeHost(ring)");
        buf.ptr = nullptr;
        buf.bytes = 0;
    }
    SCIR_CUDA(cudaHostAlloc(&buf.ptr, bytes, flags), "cudaHostAlloc(ring)");
    buf.bytes = bytes;
    return SCIR_B200_OK;
}

struct HostRegistration {          // host_stage=2: registered for the call, released on every exit path
    void* p = nullptr;
    int reg(const void* q, size_t bytes)
    {
        SCIR_CUDA(cudaHostRegister(const_cast<void*>(q), bytes, cudaHostRegisterPortable), "cudaHostRegister");
        p = const_cast<void*>(q);
        return SCIR_B200_OK;
    }
    ~HostRegistration()
    {
        if (p) cudaHostUnregister(p);
    }
};

template <typename Fn>
static int host_pipeline(scir_b200_ctx* ctx, const float* h_x, int64_t ld_x, int64_t n_in, float* h_y,
                         int64_t ld_y, int64_t n_out, int64_t batch, size_t extra_scratch_per_row, Fn&& fn)
{
    constexpr int S = scir_b200_ctx::kHostSlots;
    if (batch == 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    if (!ctx->s_h2d) {
        SCIR_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking), "cudaStreamCreate");
        SCIR_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < S; ++i) {
            SCIR_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming), "cudaEventCreate");
            SCIR_CUDA(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming), "cudaEventCreate");
            SCIR_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming), "cudaEventCreate");
        }
    }
    const size_t span_x = n_in > 0 ? (static_cast<size_t>(batch - 1) * ld_x + n_in) * 4 : 0;
    const size_t span_y = n_out > 0 ? (static_cast<size_t>(batch - 1) * ld_y + n_out) * 4 : 0;
    bool stage_x = n_in > 0 && !is_pinned_span(h_x, span_x);
    bool stage_y = n_out > 0 && !is_pinned_span(h_y, span_y);
    HostRegistration reg_x, reg_y;
    if (ctx->opt.host_stage == 0) stage_x = stage_y = false;
    if (ctx->opt.host_stage == 2 && (stage_x || stage_y)) {
        if (stage_x) SCIR_TRY(reg_x.reg(h_x, span_x));
        if (stage_y) SCIR_TRY(reg_y.reg(h_y, span_y));
        stage_x = stage_y = false;
        ctx->host_registered_calls++;
    }
    const bool staged = stage_x || stage_y;

    const int64_t ldi = (n_in + 3) / 4 * 4, ldo = (n_out + 3) / 4 * 4;
    int64_t rows = ctx->opt.host_block_rows;
    if (rows <= 0) {
        // 64 MiB of rows per block when DMA-ing caller memory directly; 16 MiB through the pinned ring, so that the
        // ring (S slots each way) stays small and a block's copy threads finish well inside one block's DMA time
        const size_t per_row = static_cast<size_t>(ldi + ldo) * 4 + extra_scratch_per_row;
        const size_t target = staged ? (size_t(16) << 20) : (size_t(64) << 20);
        rows = std::max<int64_t>(1, static_cast<int64_t>(target / std::max<size_t>(per_row, 1)));
    }
    rows = std::min(rows, batch);
    const size_t in_bytes = static_cast<size_t>(rows) * std::max<int64_t>(ldi, 4) * 4;
    const size_t out_bytes = static_cast<size_t>(rows) * std::max<int64_t>(ldo, 4) * 4;
    for (int s = 0; s < S; ++s) {
        SCIR_TRY(ctx_scratch(ctx, ctx->stage_in[s], in_bytes));
        SCIR_TRY(ctx_scratch(ctx, ctx->stage_out[s], out_bytes));
    }
    if (staged) {
        const bool wc = ctx->opt.host_stage_wc != 0;
        if (stage_x && ctx->pin_in_wc != wc)
            for (int s = 0; s < S; ++s)
                if (ctx->pin_in[s].ptr) {
                    SCIR_CUDA(cudaStreamSynchronize(ctx->s_h2d), "cudaStreamSynchronize");
                    SCIR_CUDA(cudaFreeHost(ctx->pin_in[s].ptr), "cudaFreeHost(ring)");
                    ctx->pin_in[s] = DeviceBuffer{};
                }
        ctx->pin_in_wc = wc;
        for (int s = 0; s < S; ++s) {
            if (stage_x)
                SCIR_TRY(pinned_scratch(ctx, ctx->pin_in[s], in_bytes,
                                        cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0u)));
            if (stage_y) SCIR_TRY(pinned_scratch(ctx, ctx->pin_out[s], out_bytes, cudaHostAllocPortable));
        }
        int want = static_cast<int>(ctx->opt.host_copy_threads);
        if (want <= 0) {
            // measured on the B200 boxes' 16-vCPU hosts (profiles/README.md): 4 threads 0.44, 8 threads 0.53, 12 threads
            // 0.57 of the pinned rate.  Under torchrun every local rank has its own pool: share the cores.
            int ranks = 1;
            if (const char* lw = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(lw));
            const int hw = static_cast<int>(std::thread::hardware_concurrency());
            want = std::max(2, std::min(12, (hw - 4) / ranks));
        }
        if (!ctx->pool || ctx->pool->threads() != want) {
            delete ctx->pool;
            ctx->pool = new CopyPool(want);
        }
        ctx->host_staged_calls++;
    }

    // Block b uses slot b % S.  With staging, block b's rows are copied into the pinned slot while block b-1 is on the
    // bus / in the kernel, and block b-LAG's output is copied out to the caller once its D2H has landed.
    constexpr int LAG = 3;
    const int64_t nblk = (batch + rows - 1) / rows;
    CopyPool::TicketPtr t_out[S], t_in;
    auto rows_of = [&](int64_t b) { return std::min(rows, batch - b * rows); };
    auto drain = [&](int64_t b) -> int {                    // pinned slot -> caller's y rows of block b
        const int s = static_cast<int>(b % S);
        SCIR_CUDA(cudaEventSynchronize(ctx->ev_out[s]), "cudaEventSynchronize(D2H)");
        t_out[s] = ctx->pool->submit_2d(reinterpret_cast<char*>(h_y + b * rows * ld_y), static_cast<size_t>(ld_y) * 4,
                                        static_cast<const char*>(ctx->pin_out[s].ptr), static_cast<size_t>(ldo) * 4,
                                        static_cast<size_t>(n_out) * 4, static_cast<size_t>(rows_of(b)), ctx->opt.host_stage_nt != 0);
        return SCIR_B200_OK;
    };
    auto run_blocks = [&]() -> int {
    for (int64_t blk = 0; blk < nblk; ++blk) {
        const int s = static_cast<int>(blk % S);
        const int64_t r0 = blk * rows, nr = rows_of(blk);
        float* din = static_cast<float*>(ctx->stage_in[s].ptr);
        float* dout = static_cast<float*>(ctx->stage_out[s].ptr);
        if (stage_x) {
            if (blk >= S) SCIR_CUDA(cudaEventSynchronize(ctx->ev_in[s]), "cudaEventSynchronize(H2D)");   // slot's last DMA read done
            t_in = ctx->pool->submit_2d(static_cast<char*>(ctx->pin_in[s].ptr), static_cast<size_t>(ldi) * 4,
                                        reinterpret_cast<const char*>(h_x + r0 * ld_x), static_cast<size_t>(ld_x) * 4,
                                        static_cast<size_t>(n_in) * 4, static_cast<size_t>(nr), ctx->opt.host_stage_nt != 0);
        }
        if (stage_y && blk >= LAG) SCIR_TRY(drain(blk - LAG));
        if (stage_x) ctx->pool->wait(t_in);
        if (stage_y && t_out[s]) {                          // the slot's previous contents must have reached the caller
            ctx->pool->wait(t_out[s]);
            t_out[s].reset();
        }
        if (blk >= S) SCIR_CUDA(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_k[s], 0), "cudaStreamWaitEvent");
        if (n_in > 0) {
            if (stage_x)
                SCIR_CUDA(cudaMemcpyAsync(din, ctx->pin_in[s].ptr, static_cast<size_t>(nr) * ldi * 4, cudaMemcpyHostToDevice,
                                          ctx->s_h2d),
                          "cudaMemcpyAsync(H2D)");
            else if (ld_x == n_in && ldi == n_in)           // dense rows: one linear DMA
                SCIR_CUDA(cudaMemcpyAsync(din, h_x + r0 * ld_x, static_cast<size_t>(nr) * n_in * 4, cudaMemcpyHostToDevice,
                                          ctx->s_h2d),
                          "cudaMemcpyAsync(H2D)");
            else
                SCIR_CUDA(cudaMemcpy2DAsync(din, ldi * 4, h_x + r0 * ld_x, ld_x * 4, n_in * 4, nr,
                                            cudaMemcpyHostToDevice, ctx->s_h2d),
                          "cudaMemcpy2DAsync(H2D)");
        }
        SCIR_CUDA(cudaEventRecord(ctx->ev_in[s], ctx->s_h2d), "cudaEventRecord");
        SCIR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[s], 0), "cudaStreamWaitEvent");
        if (blk >= S) SCIR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[s], 0), "cudaStreamWaitEvent");
        SCIR_TRY(fn(nr, din, ldi, dout, ldo));
        SCIR_CUDA(cudaEventRecord(ctx->ev_k[s], ctx->stream), "cudaEventRecord");
        SCIR_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[s], 0), "cudaStreamWaitEvent");
        if (n_out > 0) {
            if (stage_y)
                SCIR_CUDA(cudaMemcpyAsync(ctx->pin_out[s].ptr, dout, static_cast<size_t>(nr) * ldo * 4, cudaMemcpyDeviceToHost,
                                          ctx->s_d2h),
                          "cudaMemcpyAsync(D2H)");
            else if (ld_y == n_out && ldo == n_out)
                SCIR_CUDA(cudaMemcpyAsync(h_y + r0 * ld_y, dout, static_cast<size_t>(nr) * n_out * 4, cudaMemcpyDeviceToHost,
                                          ctx->s_d2h),
                          "cudaMemcpyAsync(D2H)");
            else
                SCIR_CUDA(cudaMemcpy2DAsync(h_y + r0 * ld_y, ld_y * 4, dout, ldo * 4, n_out * 4, nr,
                                            cudaMemcpyDeviceToHost, ctx->s_d2h),
                          "cudaMemcpy2DAsync(D2H)");
        }
        SCIR_CUDA(cudaEventRecord(ctx->ev_out[s], ctx->s_d2h), "cudaEventRecord");
    }
    if (stage_y)
        for (int64_t b = std::max<int64_t>(0, nblk - LAG); b < nblk; ++b) SCIR_TRY(drain(b));
    return SCIR_B200_OK;
    };
    const int rc = run_blocks();
    // queued copy jobs reference caller memory: they must have run before this call returns, error or not
    if (ctx->pool) {
        ctx->pool->wait(t_in);
        for (int s = 0; s < S; ++s) ctx->pool->wait(t_out[s]);
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->s_d2h), e2 = cudaStreamSynchronize(ctx->stream),
                e3 = cudaStreamSynchronize(ctx->s_h2d);
    if (rc != SCIR_B200_OK) return rc;
    SCIR_CUDA(e1, "cudaStreamSynchronize(D2H)");
    SCIR_CUDA(e2, "cudaStreamSynchronize");
    SCIR_CUDA(e3, "cudaStreamSynchronize(H2D)");
    return SCIR_B200_OK;
}

static void reorder_taps(const float* taps, int64_t k, int order, std::vector<float>& c)
{
    c.resize(static_cast<size_t>(k));
    for (int64_t d = 0; d < k; ++d)
        c[static_cast<size_t>(d)] = (order == SCIR_B200_TAPS_SCIR) ? taps[k - 1 - d] : taps[d];
}

}  // namespace scir_b200

using namespace scir_b200;

// =====================================================================================================
extern "C" {

const char* scir_b200_version(void) { return "scir-b200 0.1.0 (sm_100a)"; }

const char* scir_b200_last_error(void) { return g_err; }

int scir_b200_device_count(int* count)
{
    if (!count) return set_error(SCIR_B200_ERR_INVALID_ARG, "count is NULL");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount");
    *count = n;
    return SCIR_B200_OK;
}

static int ctx_make(int device, void* stream, bool borrow, scir_b200_ctx** out)
{
    if (!out) return set_error(SCIR_B200_ERR_INVALID_ARG, "ctx out-pointer is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount");
    if (n == 0) return set_error(SCIR_B200_ERR_NO_DEVICE, "no CUDA device present");
    if (device < 0 || device >= n)
        return set_error(SCIR_B200_ERR_NO_DEVICE, "device %d out of range (have %d)", device, n);
    DeviceScope scope(device);
    SCIR_TRY(scope.rc);
    cudaDeviceProp prop;
    SCIR_CUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10)
        return set_error(SCIR_B200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only",
                         device, prop.major, prop.minor);
    scir_b200_ctx* c = new (std::nothrow) scir_b200_ctx();
    if (!c) return set_error(SCIR_B200_ERR_OOM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    if (borrow) {
        c->stream = static_cast<cudaStream_t>(stream);
        c->owns_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return cuda_error(e, "cudaStreamCreateWithFlags");
        }
        c->owns_stream = true;
    }
    *out = c;
    return SCIR_B200_OK;
}

int scir_b200_ctx_create(int device, scir_b200_ctx** ctx) { return ctx_make(device, nullptr, false, ctx); }

int scir_b200_ctx_create_on_stream(int device, void* cuda_stream, scir_b200_ctx** ctx)
{
    return ctx_make(device, cuda_stream, true, ctx);
}

int scir_b200_ctx_destroy(scir_b200_ctx* ctx)
{
    if (!ctx) return SCIR_B200_OK;
    DeviceScope scope(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->s_h2d) cudaStreamSynchronize(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
    if (ctx->scratch.ptr) cudaFree(ctx->scratch.ptr);
    if (ctx->toep_flags.ptr) cudaFree(ctx->toep_flags.ptr);
    if (ctx->row_bg.ptr) cudaFree(ctx->row_bg.ptr);
    if (ctx->toep_taps.ptr) cudaFree(ctx->toep_taps.ptr);
    if (ctx->os_tw.ptr) cudaFree(ctx->os_tw.ptr);
    if (ctx->os_taps.ptr) cudaFree(ctx->os_taps.ptr);
    if (ctx->os_H.ptr) cudaFree(ctx->os_H.ptr);
    if (ctx->gen_taps.ptr) cudaFree(ctx->gen_taps.ptr);
    delete ctx->pool;
    for (int i = 0; i < scir_b200_ctx::kHostSlots; ++i) {
        if (ctx->pin_in[i].ptr) cudaFreeHost(ctx->pin_in[i].ptr);
        if (ctx->pin_out[i].ptr) cudaFreeHost(ctx->pin_out[i].ptr);
        if (ctx->stage_in[i].ptr) cudaFree(ctx->stage_in[i].ptr);
        if (ctx->stage_out[i].ptr) cudaFree(ctx->stage_out[i].ptr);
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SCIR_B200_OK;
}

int scir_b200_ctx_sync(scir_b200_ctx* ctx)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(ctx_bind(ctx));
    SCIR_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    return SCIR_B200_OK;
}

int scir_b200_ctx_device(const scir_b200_ctx* ctx, int* device)
{
    SCIR_TRY(check_ctx(ctx));
    if (!device) return set_error(SCIR_B200_ERR_INVALID_ARG, "device is NULL");
    *device = ctx->device;
    return SCIR_B200_OK;
}

int scir_b200_ctx_stream(const scir_b200_ctx* ctx, void** cuda_stream)
{
    SCIR_TRY(check_ctx(ctx));
    if (!cuda_stream) return set_error(SCIR_B200_ERR_INVALID_ARG, "cuda_stream is NULL");
    *cuda_stream = ctx->stream;
    return SCIR_B200_OK;
}

static int64_t* option_slot(Options& o, const char* key)
{
    if (!key) return nullptr;
    if (!strcmp(key, "variant")) return &o.variant;
    if (!strcmp(key, "host_block_rows")) return &o.host_block_rows;
    if (!strcmp(key, "long_tap_path")) return &o.long_tap_path;
    if (!strcmp(key, "upfirdn_variant")) return &o.upfirdn_variant;
    if (!strcmp(key, "upfirdn_ws_stages")) return &o.upfirdn_ws_stages;
    if (!strcmp(key, "toeplitz_terms")) return &o.toeplitz_terms;
    if (!strcmp(key, "toeplitz_split")) return &o.toeplitz_split;
    if (!strcmp(key, "toeplitz_chains")) return &o.toeplitz_chains;
    if (!strcmp(key, "toeplitz_ts")) return &o.toeplitz_ts;
    if (!strcmp(key, "toeplitz_adaptive_budget")) return &o.toeplitz_adaptive_budget;
    if (!strcmp(key, "toeplitz_stcs")) return &o.toeplitz_stcs;
    if (!strcmp(key, "toeplitz_tn")) return &o.toeplitz_tn;
    if (!strcmp(key, "toeplitz_tn_short")) return &o.toeplitz_tn_short;
    if (!strcmp(key, "ffma2")) return &o.ffma2;
    if (!strcmp(key, "filtfilt_fused")) return &o.filtfilt_fused;
    if (!strcmp(key, "toeplitz_min_k")) return &o.toeplitz_min_k;
    if (!strcmp(key, "os_min_k")) return &o.os_min_k;
    if (!strcmp(key, "os_packed")) return &o.os_packed;
    if (!strcmp(key, "toeplitz_min_k_full")) return &o.toeplitz_min_k_full;
    if (!strcmp(key, "toeplitz_loader")) return &o.toeplitz_loader;
    if (!strcmp(key, "host_stage")) return &o.host_stage;
    if (!strcmp(key, "host_copy_threads")) return &o.host_copy_threads;
    if (!strcmp(key, "host_stage_wc")) return &o.host_stage_wc;
    if (!strcmp(key, "host_stage_nt")) return &o.host_stage_nt;
    return nullptr;
}

int scir_b200_ctx_set_option(scir_b200_ctx* ctx, const char* key, int64_t value)
{
    SCIR_TRY(check_ctx(ctx));
    int64_t* s = option_slot(ctx->opt, key);
    if (!s) return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown option '%s'", key ? key : "(null)");
    *s = value;
    return SCIR_B200_OK;
}

int scir_b200_ctx_get_option(const scir_b200_ctx* ctx, const char* key, int64_t* value)
{
    SCIR_TRY(check_ctx(ctx));
    if (!value) return set_error(SCIR_B200_ERR_INVALID_ARG, "value is NULL");
    if (key && !strcmp(key, "toeplitz_launches")) {                   // read-only statistic
        *value = static_cast<int64_t>(ctx->toeplitz_launches);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "toeplitz_mma_per_tile")) {            // read-only statistic (last Toeplitz launch)
        *value = ctx->toeplitz_last_mma_per_tile;
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "toeplitz_hh_blocks")) {               // read-only statistic (last Toeplitz launch)
        *value = ctx->toeplitz_last_hh_blocks;
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "os_launches")) {                      // read-only statistic
        *value = static_cast<int64_t>(ctx->os_launches);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "fixup_launches")) {                   // read-only statistic
        *value = static_cast<int64_t>(ctx->fixup_launches);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "gen_tiled_launches")) {                   // read-only statistic
        *value = static_cast<int64_t>(ctx->gen_tiled_launches);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "filtfilt_fused_calls")) {                   // read-only statistic
        *value = static_cast<int64_t>(ctx->filtfilt_fused_calls);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "host_staged_calls")) {                   // read-only statistic
        *value = static_cast<int64_t>(ctx->host_staged_calls);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "host_registered_calls")) {               // read-only statistic
        *value = static_cast<int64_t>(ctx->host_registered_calls);
        return SCIR_B200_OK;
    }
    if (key && !strcmp(key, "poly_launches")) {                       // read-only statistic
        *value = static_cast<int64_t>(ctx->poly_launches);
        return SCIR_B200_OK;
    }
    int64_t* s = option_slot(const_cast<scir_b200_ctx*>(ctx)->opt, key);
    if (!s) return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown option '%s'", key ? key : "(null)");
    *value = *s;
    return SCIR_B200_OK;
}

int scir_b200_ctx_launch_count(const scir_b200_ctx* ctx, uint64_t* count)
{
    SCIR_TRY(check_ctx(ctx));
    if (!count) return set_error(SCIR_B200_ERR_INVALID_ARG, "count is NULL");
    *count = ctx->launches;
    return SCIR_B200_OK;
}

// ---- memory ------------------------------------------------------------------------------------------
int scir_b200_malloc(scir_b200_ctx* ctx, size_t bytes, void** d_ptr)
{
    SCIR_ENTER(ctx);
    if (!d_ptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "d_ptr is NULL");
    *d_ptr = nullptr;
    if (bytes == 0) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    SCIR_CUDA(cudaMalloc(d_ptr, bytes), "cudaMalloc");
    return SCIR_B200_OK;
}

int scir_b200_free(scir_b200_ctx* ctx, void* d_ptr)
{
    SCIR_ENTER(ctx);
    if (!d_ptr) return SCIR_B200_OK;
    SCIR_TRY(ctx_bind(ctx));
    SCIR_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    SCIR_CUDA(cudaFree(d_ptr), "cudaFree");
    return SCIR_B200_OK;
}

int scir_b200_memcpy_h2d(scir_b200_ctx* ctx, void* d_dst, const void* h_src, size_t bytes)
{
    SCIR_ENTER(ctx);
    if (bytes == 0) return SCIR_B200_OK;
    if (!d_dst || !h_src) return set_error(SCIR_B200_ERR_INVALID_ARG, "memcpy_h2d: NULL pointer");
    SCIR_TRY(ctx_bind(ctx));
    SCIR_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(H2D)");
    SCIR_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    return SCIR_B200_OK;
}

int scir_b200_memcpy_d2h(scir_b200_ctx* ctx, void* h_dst, const void* d_src, size_t bytes)
{
    SCIR_ENTER(ctx);
    if (bytes == 0) return SCIR_B200_OK;
    if (!h_dst || !d_src) return set_error(SCIR_B200_ERR_INVALID_ARG, "memcpy_d2h: NULL pointer");
    SCIR_TRY(ctx_bind(ctx));
    SCIR_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(D2H)");
    SCIR_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    return SCIR_B200_OK;
}

int scir_b200_host_alloc(size_t bytes, void** h_ptr)
{
    if (!h_ptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "h_ptr is NULL");
    *h_ptr = nullptr;
    if (bytes == 0) return SCIR_B200_OK;
    SCIR_CUDA(cudaHostAlloc(h_ptr, bytes, cudaHostAllocPortable), "cudaHostAlloc");
    return SCIR_B200_OK;
}

int scir_b200_host_free(void* h_ptr)
{
    if (!h_ptr) return SCIR_B200_OK;
    SCIR_CUDA(cudaFreeHost(h_ptr), "cudaFreeHost");
    return SCIR_B200_OK;
}

int scir_b200_host_register(void* h_ptr, size_t bytes)
{
    if (!h_ptr || bytes == 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "host_register: NULL pointer or zero size");
    cudaError_t e = cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return SCIR_B200_OK;
    }
    SCIR_CUDA(e, "cudaHostRegister");
    return SCIR_B200_OK;
}

int scir_b200_host_unregister(void* h_ptr)
{
    if (!h_ptr) return SCIR_B200_OK;
    SCIR_CUDA(cudaHostUnregister(h_ptr), "cudaHostUnregister");
    return SCIR_B200_OK;
}

int scir_b200_host_is_pinned(const void* h_ptr, size_t bytes, int* pinned)
{
    if (!pinned) return set_error(SCIR_B200_ERR_INVALID_ARG, "pinned is NULL");
    *pinned = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount");
    if (n == 0) return set_error(SCIR_B200_ERR_NO_DEVICE, "no CUDA device present");
    *pinned = (h_ptr != nullptr && bytes > 0 && is_pinned_span(h_ptr, bytes)) ? 1 : 0;
    return SCIR_B200_OK;
}

int scir_b200_current_device(int* device)
{
    if (!device) return set_error(SCIR_B200_ERR_INVALID_ARG, "device is NULL");
    *device = -1;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount");
    if (n == 0) return set_error(SCIR_B200_ERR_NO_DEVICE, "no CUDA device present");
    SCIR_CUDA(cudaGetDevice(device), "cudaGetDevice");
    return SCIR_B200_OK;
}

// ---- hot path ------------------------------------------------------------------------------------------
int scir_b200_fir1d_batched_f32(scir_b200_ctx* ctx, const float* d_x, int64_t ld_x, const float* taps, int64_t k,
                                int tap_order, float* d_y, int64_t ld_y, int64_t batch, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_taps(taps, k));
    if (tap_order != SCIR_B200_TAPS_SCIR && tap_order != SCIR_B200_TAPS_LFILTER)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown tap_order %d", tap_order);
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix(d_y, ld_y, batch, n, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n, d_y, ld_y, n, batch));
    std::vector<float> c;
    reorder_taps(taps, k, tap_order, c);
    return fir_causal(ctx, c.data(), k, d_x, ld_x, d_y, ld_y, batch, n);
}

int scir_b200_fir1d_batched_f32_host(scir_b200_ctx* ctx, const float* h_x, int64_t ld_x, const float* taps,
                                     int64_t k, int tap_order, float* h_y, int64_t ld_y, int64_t batch, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_taps(taps, k));
    if (tap_order != SCIR_B200_TAPS_SCIR && tap_order != SCIR_B200_TAPS_LFILTER)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown tap_order %d", tap_order);
    SCIR_TRY(check_matrix(h_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix(h_y, ld_y, batch, n, "y"));
    if (n == 0) return SCIR_B200_OK;
    std::vector<float> c;
    reorder_taps(taps, k, tap_order, c);
    return host_pipeline(ctx, h_x, ld_x, n, h_y, ld_y, n, batch, 0,
                         [&](int64_t nr, const float* din, int64_t ldi, float* dout, int64_t ldo) {
                             return fir_causal(ctx, c.data(), k, din, ldi, dout, ldo, nr, n);
                         });
}

int scir_b200_lfilter_fir_f32(scir_b200_ctx* ctx, const float* b, int64_t k, float a0, const float* d_x,
                              int64_t ld_x, const float* d_zi, float* d_zf, float* d_y, int64_t ld_y,
                              int64_t batch, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_taps(b, k));
    if (a0 == 0.f) return set_error(SCIR_B200_ERR_INVALID_ARG, "a[0] must be nonzero");
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix(d_y, ld_y, batch, n, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n, d_y, ld_y, n, batch));
    std::vector<float> c(static_cast<size_t>(k));
    for (int64_t d = 0; d < k; ++d) c[static_cast<size_t>(d)] = b[d] / a0;     // _signaltools.py:2223
    SCIR_TRY(fir_causal(ctx, c.data(), k, d_x, ld_x, d_y, ld_y, batch, n));
    if (d_zi) SCIR_TRY(launch_add_zi(ctx, d_y, ld_y, d_zi, batch, n, k));
    if (d_zf) SCIR_TRY(launch_compute_zf(ctx, c.data(), k, d_x, ld_x, d_zi, d_zf, batch, n));
    return SCIR_B200_OK;
}

int64_t scir_b200_upfirdn_out_len(int64_t len_h, int64_t in_len, int64_t up, int64_t down)
{
    return upfirdn_out_len(len_h, in_len, up, down);
}

int scir_b200_upfirdn_f32(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                          const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y, int64_t ld_y,
                          int64_t m_begin, int64_t m_count)
{
    SCIR_ENTER(ctx);
    if (!h || len_h < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "h must hold at least one tap");
    if (up < 1 || down < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "up and down must be >= 1");   // _upfirdn.py:98
    if (n_in < 1 && batch > 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "upfirdn needs n_in >= 1");
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n_in, "x"));
    if (m_begin < 0 || m_count < 0 || (batch > 0 && m_begin + m_count > upfirdn_out_len(len_h, n_in, up, down)))
        return set_error(SCIR_B200_ERR_INVALID_ARG, "output window [%lld, %lld) outside the upfirdn result",
                         (long long)m_begin, (long long)(m_begin + m_count));
    SCIR_TRY(check_matrix(d_y, ld_y, batch, m_count, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n_in, d_y, ld_y, m_count, batch));
    if (batch == 0 || m_count == 0) return SCIR_B200_OK;
    return launch_upfirdn(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count);
}

int scir_b200_resample_poly_plan(int64_t n_in, int64_t len_h, int64_t up, int64_t down,
                                 scir_b200_resample_plan* plan)
{
    if (!plan) return set_error(SCIR_B200_ERR_INVALID_ARG, "plan is NULL");
    if (up < 1 || down < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "up and down must be >= 1");
    if (n_in < 1 || len_h < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "n_in and len_h must be >= 1");
    return resample_plan(n_in, len_h, up, down, plan);
}

int scir_b200_resample_poly_f32(scir_b200_ctx* ctx, const float* window, int64_t len_h, int64_t up, int64_t down,
                                const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y,
                                int64_t ld_y)
{
    SCIR_ENTER(ctx);
    scir_b200_resample_plan pl;
    if (!window) return set_error(SCIR_B200_ERR_INVALID_ARG, "window is NULL");
    SCIR_TRY(scir_b200_resample_poly_plan(n_in, len_h, up, down, &pl));
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n_in, "x"));
    const int64_t n_out = (pl.up == 1 && pl.down == 1) ? n_in : pl.n_out;
    SCIR_TRY(check_matrix(d_y, ld_y, batch, n_out, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n_in, d_y, ld_y, n_out, batch));
    return resample_device(ctx, window, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y);
}

int scir_b200_upfirdn_mode_f32(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down, int mode,
                               float cval, const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y,
                               int64_t ld_y, int64_t m_begin, int64_t m_count)
{
    SCIR_ENTER(ctx);
    if (!h || len_h < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "h must hold at least one tap");
    if (up < 1 || down < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "up and down must be >= 1");   // _upfirdn.py:98
    if (n_in < 1 && batch > 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "upfirdn needs n_in >= 1");
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n_in, "x"));
    if (m_begin < 0 || m_count < 0 || (batch > 0 && m_begin + m_count > upfirdn_out_len(len_h, n_in, up, down)))
        return set_error(SCIR_B200_ERR_INVALID_ARG, "output window [%lld, %lld) outside the upfirdn result",
                         (long long)m_begin, (long long)(m_begin + m_count));
    SCIR_TRY(check_matrix(d_y, ld_y, batch, m_count, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n_in, d_y, ld_y, m_count, batch));
    if (batch == 0 || m_count == 0) return SCIR_B200_OK;
    return launch_upfirdn(ctx, h, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, m_begin, m_count, mode, cval);
}

int scir_b200_resample_poly_pad_f32(scir_b200_ctx* ctx, const float* window, int64_t len_h, int64_t up, int64_t down,
                                    int padtype, float cval, const float* d_x, int64_t ld_x, int64_t batch,
                                    int64_t n_in, float* d_y, int64_t ld_y)
{
    SCIR_ENTER(ctx);
    scir_b200_resample_plan pl;
    if (!window) return set_error(SCIR_B200_ERR_INVALID_ARG, "window is NULL");
    SCIR_TRY(scir_b200_resample_poly_plan(n_in, len_h, up, down, &pl));
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n_in, "x"));
    const bool copy = (pl.up == 1 && pl.down == 1);
    const int64_t n_out = copy ? n_in : pl.n_out;
    SCIR_TRY(check_matrix(d_y, ld_y, batch, n_out, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n_in, d_y, ld_y, n_out, batch));
    if (padtype >= SCIR_B200_EXT_CONSTANT && padtype <= SCIR_B200_EXT_LINE)
        return resample_device(ctx, window, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y, padtype, cval);
    if (padtype != SCIR_B200_PAD_STAT_MEAN && padtype != SCIR_B200_PAD_STAT_MINIMUM && padtype != SCIR_B200_PAD_STAT_MAXIMUM &&
        padtype != SCIR_B200_PAD_STAT_MEDIAN)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown padtype %d", padtype);
    if (batch == 0 || n_in == 0 || copy)          // SciPy returns x.copy() before looking at padtype (:3885-3886)
        return resample_device(ctx, window, len_h, up, down, d_x, ld_x, batch, n_in, d_y, ld_y);
    // background statistic per row, x - bg into scratch, zero-padded upfirdn, + bg (:3927-3957)
    const int stat = (padtype == SCIR_B200_PAD_STAT_MEAN) ? 0 : (padtype == SCIR_B200_PAD_STAT_MINIMUM) ? 1
                   : (padtype == SCIR_B200_PAD_STAT_MAXIMUM) ? 2 : 3;
    const int64_t ldc = (n_in + 3) / 4 * 4;
    SCIR_TRY(ctx_scratch(ctx, ctx->row_bg, static_cast<size_t>(batch) * sizeof(float)));
    SCIR_TRY(ctx_scratch(ctx, ctx->scratch, static_cast<size_t>(batch) * ldc * sizeof(float)));
    float* bg = static_cast<float*>(ctx->row_bg.ptr);
    float* xc = static_cast<float*>(ctx->scratch.ptr);
    SCIR_TRY(launch_row_stat(ctx, stat, d_x, ld_x, batch, n_in, bg));
    for (int64_t r0 = 0; r0 < batch; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, batch - r0);
        SCIR_TRY(launch_row_offset(ctx, d_x + r0 * ld_x, ld_x, bg + r0, -1.f, xc + r0 * ldc, ldc, nr, n_in));
    }
    SCIR_TRY(resample_device(ctx, window, len_h, up, down, xc, ldc, batch, n_in, d_y, ld_y));
    for (int64_t r0 = 0; r0 < batch; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, batch - r0);
        SCIR_TRY(launch_row_offset(ctx, d_y + r0 * ld_y, ld_y, bg + r0, +1.f, d_y + r0 * ld_y, ld_y, nr, n_out));
    }
    return SCIR_B200_OK;
}

int scir_b200_resample_poly_f32_host(scir_b200_ctx* ctx, const float* window, int64_t len_h, int64_t up,
                                     int64_t down, const float* h_x, int64_t ld_x, int64_t batch, int64_t n_in,
                                     float* h_y, int64_t ld_y)
{
    SCIR_ENTER(ctx);
    scir_b200_resample_plan pl;
    if (!window) return set_error(SCIR_B200_ERR_INVALID_ARG, "window is NULL");
    SCIR_TRY(scir_b200_resample_poly_plan(n_in, len_h, up, down, &pl));
    const int64_t n_out = (pl.up == 1 && pl.down == 1) ? n_in : pl.n_out;
    SCIR_TRY(check_matrix(h_x, ld_x, batch, n_in, "x"));
    SCIR_TRY(check_matrix(h_y, ld_y, batch, n_out, "y"));
    return host_pipeline(ctx, h_x, ld_x, n_in, h_y, ld_y, n_out, batch, 0,
                         [&](int64_t nr, const float* din, int64_t ldi, float* dout, int64_t ldo) {
                             return resample_device(ctx, window, len_h, up, down, din, ldi, nr, n_in, dout, ldo);
                         });
}

int scir_b200_filtfilt_fir_f32(scir_b200_ctx* ctx, const float* b, int64_t k, int pad_mode, int64_t padlen,
                               const float* d_x, int64_t ld_x, float* d_y, int64_t ld_y, int64_t batch, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_taps(b, k));
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix(d_y, ld_y, batch, n, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n, d_y, ld_y, n, batch));
    return filtfilt_device(ctx, b, k, pad_mode, padlen, d_x, ld_x, d_y, ld_y, batch, n);
}

int scir_b200_filtfilt_fir_f32_host(scir_b200_ctx* ctx, const float* b, int64_t k, int pad_mode, int64_t padlen,
                                    const float* h_x, int64_t ld_x, float* h_y, int64_t ld_y, int64_t batch,
                                    int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_taps(b, k));
    SCIR_TRY(check_matrix(h_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix(h_y, ld_y, batch, n, "y"));
    const size_t extra = static_cast<size_t>(n + 6 * k + 8) * 4;      // the intermediate row
    return host_pipeline(ctx, h_x, ld_x, n, h_y, ld_y, n, batch, extra,
                         [&](int64_t nr, const float* din, int64_t ldi, float* dout, int64_t ldo) {
                             return filtfilt_device(ctx, b, k, pad_mode, padlen, din, ldi, dout, ldo, nr, n);
                         });
}

int scir_b200_fir1d_batched_f64(scir_b200_ctx* ctx, const double* d_x, int64_t ld_x, const double* taps, int64_t k,
                                int tap_order, double* d_y, int64_t ld_y, int64_t batch, int64_t n)
{
    SCIR_ENTER(ctx);
    if (taps == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "taps is NULL");
    if (k < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "need at least one tap (k=%lld)", (long long)k);
    if (tap_order != SCIR_B200_TAPS_SCIR && tap_order != SCIR_B200_TAPS_LFILTER)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "unknown tap_order %d", tap_order);
    SCIR_TRY(check_matrix(d_x, ld_x, batch, n, "x"));
    SCIR_TRY(check_matrix(d_y, ld_y, batch, n, "y"));
    SCIR_TRY(check_no_alias(d_x, ld_x, n, d_y, ld_y, n, batch));
    std::vector<double> c(static_cast<size_t>(k));
    for (int64_t d = 0; d < k; ++d) c[static_cast<size_t>(d)] = (tap_order == SCIR_B200_TAPS_SCIR) ? taps[k - 1 - d] : taps[d];
    return launch_fir_f64(ctx, d_x, ld_x, c.data(), k, d_y, ld_y, batch, n);
}

// ---- DeviceArray elementwise ops on device-resident f32 (lib.rs:268-377 dispatch, :840-1034 CUDA wrappers) ----
static int check_vec(const void* p, int64_t n, const char* name)
{
    if (n < 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s: negative length", name);
    if (n > 0 && p == nullptr) return set_error(SCIR_B200_ERR_INVALID_ARG, "%s is NULL", name);
    return SCIR_B200_OK;
}

int scir_b200_add_scalar_f32(scir_b200_ctx* ctx, const float* d_a, float alpha, float* d_y, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_vec(d_a, n, "a"));
    SCIR_TRY(check_vec(d_y, n, "y"));
    return launch_elementwise(ctx, 0, d_a, nullptr, alpha, d_y, n);
}

int scir_b200_mul_scalar_f32(scir_b200_ctx* ctx, const float* d_a, float alpha, float* d_y, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_vec(d_a, n, "a"));
    SCIR_TRY(check_vec(d_y, n, "y"));
    return launch_elementwise(ctx, 1, d_a, nullptr, alpha, d_y, n);
}

int scir_b200_add_f32(scir_b200_ctx* ctx, const float* d_a, const float* d_b, float* d_y, int64_t n)
{
    SCIR_ENTER(ctx);
    SCIR_TRY(check_vec(d_a, n, "a"));
    SCIR_TRY(check_vec(d_b, n, "b"));
    SCIR_TRY(check_vec(d_y, n, "y"));
    return launch_elementwise(ctx, 2, d_a, d_b, 0.f, d_y, n);
}

int scir_b200_shard_rows(int64_t batch, int world, int rank, int64_t* row_begin, int64_t* row_end)
{
    if (!row_begin || !row_end) return set_error(SCIR_B200_ERR_INVALID_ARG, "row_begin/row_end is NULL");
    if (batch < 0 || world < 1 || rank < 0 || rank >= world)
        return set_error(SCIR_B200_ERR_INVALID_ARG, "bad shard request (batch=%lld world=%d rank=%d)",
                         (long long)batch, world, rank);
    const int64_t base = batch / world, rem = batch % world;
    *row_begin = rank * base + std::min<int64_t>(rank, rem);
    *row_end = *row_begin + base + (rank < rem ? 1 : 0);
    return SCIR_B200_OK;
}

}  // extern "C"
