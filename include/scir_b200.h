/*
 * scir_b200.h -- C ABI of the B200-native batched-FIR backend for SciR.
 *
 * This is the drop-in boundary.  The reference reaches its GPU path only
 * through Rust functions and a hand-declared `extern "C"` block over libcuda
 * (crates/scir-gpu/src/lib.rs:549-581); this header is what a replacement `scir-gpu` crate binds
 * instead (see INTEGRATION.md for the Rust `extern "C"` block and the ctypes stub).
 *
 * Conventions
 *   - Every entry point returns `int`: 0 on success, a negative SCIR_B200_ERR_* otherwise, and
 *     records a thread-local message readable with scir_b200_last_error().  There is NO CPU
 *     fallback anywhere: no device => SCIR_B200_ERR_NO_DEVICE (the reference silently fell back,
 *     lib.rs:520-523; north_star forbids that).
 *   - Arrays are row-major (batch, n) f32 like ndarray::Array2<f32> in standard layout
 *     (lib.rs:1040-1044).  `ld_*` is the row pitch in ELEMENTS (>= row length).  All sizes are
 *     64-bit (the reference truncated b*n to u32, lib.rs:1083).
 *   - `d_*` pointers are device pointers on the ctx's device; `h_*` / taps are host pointers.
 *     Taps are always passed from the host (<= SCIR_B200_MAX_TAPS floats).  The FP32 kernels receive them as
 *     launch parameters, the tcgen05 kernel from a ctx-owned device buffer that is re-uploaded only when the
 *     filter changes; either way there is no global constant-memory state and ctxs are independent.
 *   - No entry point changes the calling thread's current CUDA device / context: the ctx's device is bound
 *     for the duration of the call and the caller's context is restored on return.
 *   - Device-pointer entry points are asynchronous on the ctx's stream; the *_host entry points
 *     (H2D + kernels + D2H, what the reference-shaped Rust functions call) return when the
 *     output is complete in host memory.
 *   - A ctx is not thread-safe; use one per thread (or lock).  Distinct ctxs are independent.
 */
#ifndef SCIR_B200_H
#define SCIR_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SCIR_B200_API __attribute__((visibility("default")))
#else
#define SCIR_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SCIR_B200_OK                 0
#define SCIR_B200_ERR_INVALID_ARG   -1   /* null pointer, negative size, bad enum, ld < n ...      */
#define SCIR_B200_ERR_NO_DEVICE     -2   /* no CUDA device / driver: maps to GpuError::BackendUnavailable */
#define SCIR_B200_ERR_OOM           -3   /* device or pinned-host allocation failed                 */
#define SCIR_B200_ERR_LAUNCH        -4   /* kernel launch / execution / copy failed                 */
#define SCIR_B200_ERR_SHAPE         -5   /* maps to GpuError::ShapeMismatch (lib.rs:62)             */
#define SCIR_B200_ERR_UNSUPPORTED   -6   /* valid request this build cannot serve (e.g. k too big)  */

#define SCIR_B200_MAX_TAPS        7936   /* the FP32 kernels' taps ride in the 32 KiB kernel-parameter space */

/* Tap order for scir_b200_fir1d_batched_f32. */
#define SCIR_B200_TAPS_SCIR          0   /* reference order: y[i] = sum_t taps[k-1-t] * x[i-t]  (lib.rs:1141-1148) */
#define SCIR_B200_TAPS_LFILTER       1   /* SciPy lfilter order: y[i] = sum_d b[d] * x[i-d]                          */

/* filtfilt edge handling. */
#define SCIR_B200_PAD_ZERO_STATE     0   /* reference structure: zero-state fwd, reverse, zero-state fwd, reverse,
                                            no extension (crates/scir-signal/src/lib.rs:278-291)                   */
#define SCIR_B200_PAD_ODD            1   /* SciPy default (scipy/signal/_signaltools.py:4745-4826)                 */
#define SCIR_B200_PAD_EVEN           2
#define SCIR_B200_PAD_CONSTANT       3
#define SCIR_B200_PAD_SCIPY_NONE     4   /* SciPy padtype=None: no extension, but lfilter_zi steady-state
                                            initial conditions are still applied (:4763-4779)                      */

typedef struct scir_b200_ctx scir_b200_ctx;   /* one device + one stream + scratch */
typedef struct scir_b200_mg  scir_b200_mg;    /* several ctxs, rows sharded across them */

/* ---- library / device ------------------------------------------------------------------------ */
SCIR_B200_API const char *scir_b200_version(void);
SCIR_B200_API const char *scir_b200_last_error(void);            /* thread-local; never NULL */
SCIR_B200_API int scir_b200_device_count(int *count);            /* replaces cuInit+cuDeviceGet probing, lib.rs:605-613 */

/* ---- context: replaces the per-call CudaCtx (lib.rs:601-622) with a long-lived handle -------- */
SCIR_B200_API int scir_b200_ctx_create(int device, scir_b200_ctx **ctx);                 /* owns a non-blocking stream */
SCIR_B200_API int scir_b200_ctx_create_on_stream(int device, void *cuda_stream,          /* borrows a cudaStream_t     */
                                   scir_b200_ctx **ctx);
SCIR_B200_API int scir_b200_ctx_destroy(scir_b200_ctx *ctx);
SCIR_B200_API int scir_b200_ctx_sync(scir_b200_ctx *ctx);                                /* replaces cuCtxSynchronize, lib.rs:1102 */
SCIR_B200_API int scir_b200_ctx_device(const scir_b200_ctx *ctx, int *device);
SCIR_B200_API int scir_b200_ctx_stream(const scir_b200_ctx *ctx, void **cuda_stream);
/* Tuning / A-B switches for profiling ("variant", "tile_rows", ...); unknown key => INVALID_ARG. */
SCIR_B200_API int scir_b200_ctx_set_option(scir_b200_ctx *ctx, const char *key, int64_t value);
SCIR_B200_API int scir_b200_ctx_get_option(const scir_b200_ctx *ctx, const char *key, int64_t *value);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
SCIR_B200_API int scir_b200_ctx_launch_count(const scir_b200_ctx *ctx, uint64_t *count);

/* ---- memory: replaces cuMemAlloc/cuMemFree/cuMemcpyHtoD/DtoH (lib.rs:1056-1066,1103-1110) ----- */
SCIR_B200_API int scir_b200_malloc(scir_b200_ctx *ctx, size_t bytes, void **d_ptr);
SCIR_B200_API int scir_b200_free(scir_b200_ctx *ctx, void *d_ptr);
SCIR_B200_API int scir_b200_memcpy_h2d(scir_b200_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);   /* sync */
SCIR_B200_API int scir_b200_memcpy_d2h(scir_b200_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);   /* sync */
SCIR_B200_API int scir_b200_host_alloc(size_t bytes, void **h_ptr);      /* pinned host memory for the *_host paths */
SCIR_B200_API int scir_b200_host_free(void *h_ptr);
/* Pin (page-lock) memory the caller already owns -- a Rust Vec / ndarray / numpy buffer that is reused across calls --
 * so the *_host entry points DMA it directly.  Unregister before freeing it.  Registering costs about as much as
 * one pass over the bytes, so it pays for buffers that live longer than one call. */
SCIR_B200_API int scir_b200_host_register(void *h_ptr, size_t bytes);
SCIR_B200_API int scir_b200_host_unregister(void *h_ptr);
/* *pinned = 1 if [h_ptr, h_ptr+bytes) is page-locked memory CUDA knows (host_alloc / host_register), else 0. */
SCIR_B200_API int scir_b200_host_is_pinned(const void *h_ptr, size_t bytes, int *pinned);
/* The calling thread's current CUDA device (what a default ctx should be created on in a multi-GPU process). */
SCIR_B200_API int scir_b200_current_device(int *device);

/* ---- the hot path: replaces fir1d_batched_f32_cuda + PTX entry (lib.rs:1036-1113, 727-811) ----
 * y[b,i] = sum_{t=0}^{min(i,k-1)} c[t] * x[b,i-t], zero initial state, same shape as x,
 * with c[t] = taps[k-1-t] (TAPS_SCIR) or taps[t] (TAPS_LFILTER).  k >= 1; batch, n >= 0. */
SCIR_B200_API int scir_b200_fir1d_batched_f32(scir_b200_ctx *ctx,
                                const float *d_x, int64_t ld_x,
                                const float *taps, int64_t k, int tap_order,
                                float *d_y, int64_t ld_y,
                                int64_t batch, int64_t n);
/* Same with host arrays (what fir1d_batched_f32_auto(.., Device::Cuda) calls, lib.rs:515-531):
 * rows are streamed through the device in blocks, H2D / kernel / D2H overlapped.  Pinned arrays
 * (scir_b200_host_alloc / scir_b200_host_register) are DMA'd in place; pageable arrays -- a Rust Vec, an
 * ndarray::Array2, a numpy buffer -- go through the ctx's pinned ring, filled and drained by its copy threads
 * (ctx option "host_stage": 1 ring [default], 0 plain cudaMemcpyAsync, 2 register the spans for the call). */
SCIR_B200_API int scir_b200_fir1d_batched_f32_host(scir_b200_ctx *ctx,
                                     const float *h_x, int64_t ld_x,
                                     const float *taps, int64_t k, int tap_order,
                                     float *h_y, int64_t ld_y,
                                     int64_t batch, int64_t n);

/* f64 twin on device-resident data: fir1d_batched_f64 (lib.rs:1166-1184; the reference has it on the CPU only,
 * as the high-precision companion its tests compare with, :1263-1298).  Same definition, IEEE f64 products and
 * sums (DFMA), newest sample first.  k <= ~23000 taps (shared-memory tile). */
SCIR_B200_API int scir_b200_fir1d_batched_f64(scir_b200_ctx *ctx,
                                const double *d_x, int64_t ld_x,
                                const double *taps, int64_t k, int tap_order,
                                double *d_y, int64_t ld_y,
                                int64_t batch, int64_t n);

/* ---- scir-signal FIR routes that map onto the hot path (SciPy is the spec) ---- */

/* lfilter(b, [a0], x) with optional streaming state (scipy/signal/_signaltools.py:2181-2242).
 * d_zi / d_zf: (batch, k-1) row-major device arrays, or NULL (zero state in, state not returned). */
SCIR_B200_API int scir_b200_lfilter_fir_f32(scir_b200_ctx *ctx,
                              const float *b, int64_t k, float a0,
                              const float *d_x, int64_t ld_x,
                              const float *d_zi, float *d_zf,
                              float *d_y, int64_t ld_y,
                              int64_t batch, int64_t n);

/* _output_len (scipy/signal/_upfirdn_apply.pyx:59-67), pure int64. */
SCIR_B200_API int64_t scir_b200_upfirdn_out_len(int64_t len_h, int64_t in_len, int64_t up, int64_t down);

/* Signal-extension modes of upfirdn / resample_poly: values = SciPy's MODE enum (_upfirdn_apply.pyx:77-86),
 * names = its mode strings (pyx:89-105: 'constant', 'symmetric', 'edge', 'smooth', 'wrap', 'reflect',
 * 'antisymmetric', 'antireflect', 'line').  Modes that reflect need n_in >= 2 where SciPy divides by n_in-1. */
enum {
    SCIR_B200_EXT_CONSTANT = 0, SCIR_B200_EXT_SYMMETRIC = 1, SCIR_B200_EXT_EDGE = 2, SCIR_B200_EXT_SMOOTH = 3,
    SCIR_B200_EXT_PERIODIC = 4, SCIR_B200_EXT_REFLECT = 5, SCIR_B200_EXT_ANTISYMMETRIC = 6,
    SCIR_B200_EXT_ANTIREFLECT = 7, SCIR_B200_EXT_LINE = 8
};
/* resample_poly padtypes beyond the extension modes (_signaltools.py:3921-3957): a per-row statistic is removed
 * before and restored after a zero-padded upfirdn (the median is an exact per-row radix select on the device). */
enum { SCIR_B200_PAD_STAT_MEAN = 16, SCIR_B200_PAD_STAT_MEDIAN = 17, SCIR_B200_PAD_STAT_MINIMUM = 18, SCIR_B200_PAD_STAT_MAXIMUM = 19 };

/* upfirdn(h, x, up, down), mode='constant' (pyx:421-481): writes outputs m in
 * [m_begin, m_begin+m_count) of each row to d_y[b, 0..m_count); the full result is
 * m_begin=0, m_count=scir_b200_upfirdn_out_len(len_h, n_in, up, down). */
SCIR_B200_API int scir_b200_upfirdn_f32(scir_b200_ctx *ctx,
                          const float *h, int64_t len_h, int64_t up, int64_t down,
                          const float *d_x, int64_t ld_x, int64_t batch, int64_t n_in,
                          float *d_y, int64_t ld_y, int64_t m_begin, int64_t m_count);

/* Same with a signal-extension mode (pyx:110-231): samples outside [0, n_in) take the mode's value instead of 0. */
SCIR_B200_API int scir_b200_upfirdn_mode_f32(scir_b200_ctx *ctx,
                               const float *h, int64_t len_h, int64_t up, int64_t down, int mode, float cval,
                               const float *d_x, int64_t ld_x, int64_t batch, int64_t n_in,
                               float *d_y, int64_t ld_y, int64_t m_begin, int64_t m_count);

/* Integer plan of resample_poly (scipy/signal/_signaltools.py:3882-3918); bit-exact contract. */
typedef struct {
    int64_t up, down;          /* after gcd reduction                          */
    int64_t n_out;             /* ceil(n_in*up/down)                           */
    int64_t half_len;          /* (len_h-1)/2                                  */
    int64_t n_pre_pad;         /* down - half_len % down                       */
    int64_t n_post_pad;
    int64_t n_pre_remove;      /* (half_len + n_pre_pad) / down                */
    int64_t len_h_padded;      /* len_h + n_pre_pad + n_post_pad               */
    int64_t upfirdn_len;       /* _output_len(len_h_padded, n_in, up, down)    */
} scir_b200_resample_plan;
SCIR_B200_API int scir_b200_resample_poly_plan(int64_t n_in, int64_t len_h, int64_t up, int64_t down,
                                 scir_b200_resample_plan *plan);

/* resample_poly(x, up, down, window=h) with padtype='constant', cval=0.  d_y is (batch, n_out).
 * If up/gcd == down/gcd == 1 the input is copied (SciPy :3885-3886). */
SCIR_B200_API int scir_b200_resample_poly_f32(scir_b200_ctx *ctx,
                                const float *window, int64_t len_h, int64_t up, int64_t down,
                                const float *d_x, int64_t ld_x, int64_t batch, int64_t n_in,
                                float *d_y, int64_t ld_y);
/* resample_poly(..., padtype, cval): padtype = SCIR_B200_EXT_* or SCIR_B200_PAD_STAT_*; cval for EXT_CONSTANT. */
SCIR_B200_API int scir_b200_resample_poly_pad_f32(scir_b200_ctx *ctx,
                                    const float *window, int64_t len_h, int64_t up, int64_t down,
                                    int padtype, float cval,
                                    const float *d_x, int64_t ld_x, int64_t batch, int64_t n_in,
                                    float *d_y, int64_t ld_y);
SCIR_B200_API int scir_b200_resample_poly_f32_host(scir_b200_ctx *ctx,
                                     const float *window, int64_t len_h, int64_t up, int64_t down,
                                     const float *h_x, int64_t ld_x, int64_t batch, int64_t n_in,
                                     float *h_y, int64_t ld_y);

/* Zero-phase forward-backward filtering with an FIR numerator b (lfilter order), a = [1].
 * PAD_ZERO_STATE reproduces the reference's structure; ODD/EVEN/CONSTANT/SCIPY_NONE reproduce
 * SciPy's filtfilt(method='pad'); padlen < 0 means SciPy's default 3*k (ignored for ZERO_STATE and
 * SCIPY_NONE).  For the SciPy modes n must exceed the pad length (SciPy raises ValueError:
 * SCIR_B200_ERR_SHAPE here). */
SCIR_B200_API int scir_b200_filtfilt_fir_f32(scir_b200_ctx *ctx,
                               const float *b, int64_t k, int pad_mode, int64_t padlen,
                               const float *d_x, int64_t ld_x,
                               float *d_y, int64_t ld_y,
                               int64_t batch, int64_t n);
SCIR_B200_API int scir_b200_filtfilt_fir_f32_host(scir_b200_ctx *ctx,
                                    const float *b, int64_t k, int pad_mode, int64_t padlen,
                                    const float *h_x, int64_t ld_x,
                                    float *h_y, int64_t ld_y,
                                    int64_t batch, int64_t n);

/* ---- f64 twins of the scir-signal routes on device-resident data --------------------------
 * The reference's own resample_poly / filtfilt are f64 (crates/scir-signal/src/lib.rs:278-291, :313-362) and so are its
 * fixtures (:638-668); these entry points serve them in f64 (IEEE DFMA), same semantics as the f32 ones above. */
SCIR_B200_API int scir_b200_upfirdn_mode_f64(scir_b200_ctx *ctx,
                               const double *h, int64_t len_h, int64_t up, int64_t down, int mode, double cval,
                               const double *d_x, int64_t ld_x, int64_t batch, int64_t n_in,
                               double *d_y, int64_t ld_y, int64_t m_begin, int64_t m_count);
SCIR_B200_API int scir_b200_resample_poly_pad_f64(scir_b200_ctx *ctx,
                                    const double *window, int64_t len_h, int64_t up, int64_t down,
                                    int padtype, double cval,
                                    const double *d_x, int64_t ld_x, int64_t batch, int64_t n_in,
                                    double *d_y, int64_t ld_y);
SCIR_B200_API int scir_b200_filtfilt_fir_f64(scir_b200_ctx *ctx,
                               const double *b, int64_t k, int pad_mode, int64_t padlen,
                               const double *d_x, int64_t ld_x,
                               double *d_y, int64_t ld_y,
                               int64_t batch, int64_t n);

/* ---- DeviceArray<f32> elementwise ops on device-resident data -----------------
 * Replace add_scalar_f32_cuda / mul_scalar_f32_cuda / add_vec_f32_cuda (lib.rs:912-1034, 840-911), the
 * Device::Cuda arms of add_scalar_auto / mul_scalar_auto / add_auto (lib.rs:268-377).  n elements, any
 * alignment; y may alias a (and b).  Bit-identical to the CPU loops lib.rs:206-255 (one add or mul, no FMA). */
SCIR_B200_API int scir_b200_add_scalar_f32(scir_b200_ctx *ctx, const float *d_a, float alpha, float *d_y, int64_t n);
SCIR_B200_API int scir_b200_mul_scalar_f32(scir_b200_ctx *ctx, const float *d_a, float alpha, float *d_y, int64_t n);
SCIR_B200_API int scir_b200_add_f32(scir_b200_ctx *ctx, const float *d_a, const float *d_b, float *d_y, int64_t n);

/* ---- multi-GPU front end: rows (channels) sharded across devices, no collective --- */
SCIR_B200_API int scir_b200_mg_create(const int *devices, int n_devices, scir_b200_mg **mg);
SCIR_B200_API int scir_b200_mg_destroy(scir_b200_mg *mg);
SCIR_B200_API int scir_b200_mg_device_count(const scir_b200_mg *mg, int *n_devices);
/* Borrow shard `shard`'s ctx (owned by mg) for scir_b200_malloc / memcpy on that device. */
SCIR_B200_API int scir_b200_mg_ctx(const scir_b200_mg *mg, int shard, scir_b200_ctx **ctx);
SCIR_B200_API int scir_b200_mg_sync(scir_b200_mg *mg);                       /* waits for every device's stream */
/* Row block [row_begin, row_end) of a `batch`-row problem owned by shard `rank` of `world`
 * (contiguous blocks, remainder spread over the first shards).  Pure integer; shared with the
 * one-process-per-GPU launcher so both front ends shard identically. */
SCIR_B200_API int scir_b200_shard_rows(int64_t batch, int world, int rank, int64_t *row_begin, int64_t *row_end);
SCIR_B200_API int scir_b200_mg_fir1d_batched_f32_host(scir_b200_mg *mg,
                                        const float *h_x, int64_t ld_x,
                                        const float *taps, int64_t k, int tap_order,
                                        float *h_y, int64_t ld_y,
                                        int64_t batch, int64_t n);

SCIR_B200_API int scir_b200_mg_resample_poly_f32_host(scir_b200_mg *mg,
                                        const float *window, int64_t len_h, int64_t up, int64_t down,
                                        const float *h_x, int64_t ld_x, int64_t batch, int64_t n_in,
                                        float *h_y, int64_t ld_y);
SCIR_B200_API int scir_b200_mg_filtfilt_fir_f32_host(scir_b200_mg *mg,
                                       const float *b, int64_t k, int pad_mode, int64_t padlen,
                                       const float *h_x, int64_t ld_x,
                                       float *h_y, int64_t ld_y,
                                       int64_t batch, int64_t n);

/* Device-resident shards: d_x[s] / d_y[s] point at shard s's first row ON THE DEVICE OF SHARD s
 * (scir_b200_shard_rows(batch, world, s) rows, pitch ld_x[s] / ld_y[s]).  Asynchronous on every device's stream;
 * scir_b200_mg_sync() waits.  Same semantics per row as scir_b200_fir1d_batched_f32. */
SCIR_B200_API int scir_b200_mg_fir1d_batched_f32(scir_b200_mg *mg,
                                   const float *const *d_x, const int64_t *ld_x,
                                   const float *taps, int64_t k, int tap_order,
                                   float *const *d_y, const int64_t *ld_y,
                                   int64_t batch, int64_t n);
/* The optional "whole output on one device" step (north_star (d); SURVEY.md 8(e): never on the hot path): fans the
 * row shards in to d_dst (batch, n), pitch ld_dst, on the device of shard `dst_shard`.  Peer-to-peer copies over
 * NVLink, one per shard, each queued on the source device's stream behind that shard's kernel; returns when all
 * have landed. */
SCIR_B200_API int scir_b200_mg_gather_rows_f32(scir_b200_mg *mg,
                                 const float *const *d_shards, const int64_t *ld_shards,
                                 int dst_shard, float *d_dst, int64_t ld_dst,
                                 int64_t batch, int64_t n);

/* ---- measurement helpers (bench.py): what the SAME box sustains, as roofline denominators ----- */
SCIR_B200_API int scir_b200_microbench_ffma(scir_b200_ctx *ctx, int iters, double *tflops);      /* FP32 FFMA peak   */
SCIR_B200_API int scir_b200_microbench_ffma2(scir_b200_ctx *ctx, int iters, int mix, double *tflops); /* packed FFMA2 (+mix) */
SCIR_B200_API int scir_b200_microbench_copy(scir_b200_ctx *ctx, size_t bytes, int iters, double *gbps); /* HBM rd+wr */
/* PCIe ceiling of the *_host entry points: pinned copies of `bytes`, H2D alone, D2H alone, both at once (each way). */
SCIR_B200_API int scir_b200_microbench_pcie(scir_b200_ctx *ctx, size_t bytes, int iters,
                              double *h2d_gbs, double *d2h_gbs, double *duplex_each_gbs);

#ifdef __cplusplus
}
#endif
#endif /* SCIR_B200_H */
