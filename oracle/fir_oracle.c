/*
 * oracle/fir_oracle.c -- CPU restatement of the reference's batched-FIR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under scir_b200/ (the product) may link,
 * import or call this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the *checker* and as the
 * timed CPU baseline -- never as a fallback for the CUDA path.
 *
 * Build: `make -C oracle` (gcc -O2 -ffp-contract=off; contraction is disabled so
 * the f32 path rounds exactly like the reference's Rust, which never fuses
 * mul+add).
 *
 * What it restates (paths relative to /root/reference):
 *   - crates/scir-gpu/src/lib.rs:1134-1152  fir1d_batched_f32   (oracle_fir1d_batched_f32)
 *   - crates/scir-gpu/src/lib.rs:1166-1184  fir1d_batched_f64   (oracle_fir1d_batched_f64)
 *   - crates/scir-signal/src/lib.rs:293-362 convolve + legacy 2/3 resample_poly
 *   - crates/scir-signal/src/lib.rs:278-291 filtfilt structure (zero-state forward,
 *     reverse, zero-state forward, reverse) -- restated here for an FIR numerator
 *   - SciPy 1.17.0.dev0 (vendored, un-built submodule; importable scipy here is 1.18.1)
 *     for the API the reference does not have:
 *       scipy/signal/_upfirdn_apply.pyx:59-67   _output_len
 *       scipy/signal/_upfirdn_apply.pyx:421-481 _apply_impl (mode='constant', cval=0)
 *       scipy/signal/_signaltools.py:3865-3957  resample_poly (array `window`)
 *       scipy/signal/_signaltools.py:2181-2242  lfilter, FIR branch (a=[a0])
 *       scipy/signal/_signaltools.py:4745-4826  filtfilt(method='pad')
 *       scipy/signal/_arraytools.py:57-107      odd_ext
 *
 * Parity pinning: tests/test_oracle_golden.py checks every function here against
 * the reference's own known-answer vectors and against
 * installed SciPy; see DESIGN.md "Oracle".
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Reference hot loop, f32 (gpu/lib.rs:1134-1152).                            */
/* y[b,i] = sum over t=0..min(i,k-1) of taps[k-1-t] * x[b,i-t], newest first, */
/* f32 multiply then f32 add, zero initial state, same-length output.         */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_fir1d_batched_f32(const float *x, int64_t batch, int64_t n,
                                         const float *taps, int64_t k, float *y)
{
    for (int64_t b = 0; b < batch; ++b) {
        const float *xin = x + b * n;
        float *yout = y + b * n;
        for (int64_t i = 0; i < n; ++i) {
            float acc = 0.0f;
            int64_t start = (i + 1 >= k) ? (i + 1 - k) : 0;   /* saturating_sub */
            int64_t t = 0;
            for (int64_t xi = i; xi >= start; --xi, ++t) {
                float prod = taps[k - 1 - t] * xin[xi];
                acc = acc + prod;
            }
            yout[i] = acc;
        }
    }
}

/* Same loop spread over host threads by contiguous row blocks (rows are independent,
 * SURVEY 8e).  Used only for the "all cores" CPU-baseline figure; the reference itself
 * is single-threaded.  Returns the number of threads actually used. */
typedef struct {
    const float *x, *taps;
    float *y;
    int64_t r0, r1, n, k;
} fir_mt_job;

static void *fir_mt_worker(void *arg)
{
    fir_mt_job *j = (fir_mt_job *)arg;
    if (j->r1 > j->r0)
        oracle_fir1d_batched_f32(j->x + j->r0 * j->n, j->r1 - j->r0, j->n, j->taps, j->k,
                                 j->y + j->r0 * j->n);
    return NULL;
}

ORACLE_API int oracle_fir1d_batched_f32_mt(const float *x, int64_t batch, int64_t n,
                                           const float *taps, int64_t k, float *y,
                                           int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((int64_t)threads > batch) threads = (int)(batch > 0 ? batch : 1);
    pthread_t tid[256];
    fir_mt_job job[256];
    for (int t = 0; t < threads; ++t) {
        job[t].x = x; job[t].taps = taps; job[t].y = y; job[t].n = n; job[t].k = k;
        job[t].r0 = batch * t / threads;
        job[t].r1 = batch * (t + 1) / threads;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, fir_mt_worker, &job[t]);
    fir_mt_worker(&job[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    return threads;
}

/* f64 twin (gpu/lib.rs:1166-1184). */
ORACLE_API void oracle_fir1d_batched_f64(const double *x, int64_t batch, int64_t n,
                                         const double *taps, int64_t k, double *y)
{
    for (int64_t b = 0; b < batch; ++b) {
        const double *xin = x + b * n;
        double *yout = y + b * n;
        for (int64_t i = 0; i < n; ++i) {
            double acc = 0.0;
            int64_t start = (i + 1 >= k) ? (i + 1 - k) : 0;
            int64_t t = 0;
            for (int64_t xi = i; xi >= start; --xi, ++t)
                acc += taps[k - 1 - t] * xin[xi];
            yout[i] = acc;
        }
    }
}

/* High-precision judge: f32 data and taps, every product and sum in f64.
 * Tolerances (1e-5 * sum|h| * max|x|) are applied against this one, with the
 * reference-order f32 result reported beside it (hard part 3). */
ORACLE_API void oracle_fir1d_batched_f32_acc64(const float *x, int64_t batch, int64_t n,
                                               const float *taps, int64_t k, double *y)
{
    for (int64_t b = 0; b < batch; ++b) {
        const float *xin = x + b * n;
        double *yout = y + b * n;
        for (int64_t i = 0; i < n; ++i) {
            double acc = 0.0;
            int64_t start = (i + 1 >= k) ? (i + 1 - k) : 0;
            int64_t t = 0;
            for (int64_t xi = i; xi >= start; --xi, ++t)
                acc += (double)taps[k - 1 - t] * (double)xin[xi];
            yout[i] = acc;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* scir-signal legacy 2/3 resampler (sig/lib.rs:293-362), f64.                */
/* ------------------------------------------------------------------------- */

/* Full linear convolution, len n+m-1 (sig/lib.rs:293-303). */
ORACLE_API void oracle_convolve_full_f64(const double *x, int64_t n, const double *h,
                                         int64_t m, double *y)
{
    for (int64_t i = 0; i < n + m - 1; ++i) y[i] = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < m; ++j)
            y[i + j] += x[i] * h[j];
}

/* The reference bakes 31 literals (sig/lib.rs:315-347).  SURVEY 0.4 probed them to be
 * 2*firwin(31, 1/3, window='hamming') to 5.6e-17; we regenerate them by that formula
 * (Hamming-windowed sinc, unit DC gain, times up=2) instead of copying the table.
 * tests/golden/legacy_resample_taps.npy holds the reference's literals as data and
 * tests/test_oracle_golden.py pins this function against them. */
ORACLE_API void oracle_legacy_resample_taps(double *h31)
{
    const int ntaps = 31;
    const double cutoff = 1.0 / 3.0;           /* relative to Nyquist */
    const double alpha = 0.5 * (ntaps - 1);
    double s = 0.0;
    for (int i = 0; i < ntaps; ++i) {
        double m = (double)i - alpha;
        double a = M_PI * cutoff * m;
        double sinc = (m == 0.0) ? 1.0 : sin(a) / a;
        double win = 0.54 - 0.46 * cos(2.0 * M_PI * (double)i / (double)(ntaps - 1));
        h31[i] = cutoff * sinc * win;
        s += h31[i];
    }
    for (int i = 0; i < ntaps; ++i) h31[i] = 2.0 * h31[i] / s;
}

/* resample_poly(x, 2, 3) exactly as the reference does it: zero-stuff, full
 * convolution, pick every 3rd sample from offset 15 while idx < len-15
 * (sig/lib.rs:348-361).  `h` is the 31-tap filter.  Returns the output count;
 * y must hold ceil(2n/3)+1 values. */
ORACLE_API int64_t oracle_legacy_resample_poly_2_3(const double *x, int64_t n,
                                                   const double *h, double *y)
{
    const int64_t up = 2, down = 3, m = 31;
    double *stuffed = (double *)calloc((size_t)(n * up), sizeof(double));
    double *conv = (double *)malloc((size_t)(n * up + m - 1) * sizeof(double));
    for (int64_t i = 0; i < n; ++i) stuffed[i * up] = x[i];
    oracle_convolve_full_f64(stuffed, n * up, h, m, conv);
    const int64_t offset = (m - 1) / 2;
    const int64_t end = (n * up + m - 1) - offset;
    int64_t cnt = 0;
    for (int64_t idx = offset; idx < end; idx += down) y[cnt++] = conv[idx];
    free(stuffed);
    free(conv);
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* Forward-backward with an FIR numerator, reference structure                */
/* (sig/lib.rs:278-291): zero-state forward pass, zero-state pass over the    */
/* reversed intermediate, reverse.  No padding.  b is in lfilter order        */
/* (b[0] multiplies the newest sample).  f64 accumulate, f32 or f64 storage.  */
/* ------------------------------------------------------------------------- */
static void lfilter_fir_zero_state_f64(const double *b, int64_t k, const double *x,
                                       int64_t n, double *y)
{
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        int64_t dmax = (i < k - 1) ? i : (k - 1);
        for (int64_t d = 0; d <= dmax; ++d) acc += b[d] * x[i - d];
        y[i] = acc;
    }
}

ORACLE_API void oracle_filtfilt_fir_nopad_f64(const double *b, int64_t k, const double *x,
                                              int64_t batch, int64_t n, double *y)
{
    double *tmp = (double *)malloc((size_t)n * sizeof(double));
    double *rev = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t r = 0; r < batch; ++r) {
        lfilter_fir_zero_state_f64(b, k, x + r * n, n, tmp);
        for (int64_t i = 0; i < n; ++i) rev[i] = tmp[n - 1 - i];
        lfilter_fir_zero_state_f64(b, k, rev, n, tmp);
        for (int64_t i = 0; i < n; ++i) y[r * n + i] = tmp[n - 1 - i];
    }
    free(tmp);
    free(rev);
}

/* f32 storage variant: intermediate and output rounded to f32 like a two-pass f32
 * pipeline would; sums in f64 so it can serve as the judge. */
ORACLE_API void oracle_filtfilt_fir_nopad_f32(const float *b, int64_t k, const float *x,
                                              int64_t batch, int64_t n, double *y)
{
    double *bd = (double *)malloc((size_t)k * sizeof(double));
    double *xd = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t i = 0; i < k; ++i) bd[i] = b[i];
    for (int64_t r = 0; r < batch; ++r) {
        for (int64_t i = 0; i < n; ++i) xd[i] = x[r * n + i];
        oracle_filtfilt_fir_nopad_f64(bd, k, xd, 1, n, y + r * n);
    }
    free(bd);
    free(xd);
}

/* ------------------------------------------------------------------------- */
/* SciPy lfilter, FIR branch (_signaltools.py:2181-2242): b /= a0;            */
/* full = convolve(b, x); full[:k-1] += zi; y = full[:n]; zf = full[n:].       */
/* zi / zf may be NULL.  f64 arithmetic on f32 data.                          */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_lfilter_fir_f32(const float *b, int64_t k, double a0, const float *x,
                                       int64_t batch, int64_t n, const float *zi,
                                       double *y, double *zf)
{
    const int64_t full_len = n + k - 1;
    double *full = (double *)malloc((size_t)(full_len > 0 ? full_len : 1) * sizeof(double));
    for (int64_t r = 0; r < batch; ++r) {
        for (int64_t i = 0; i < full_len; ++i) full[i] = 0.0;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t d = 0; d < k; ++d)
                full[i + d] += ((double)b[d] / a0) * (double)x[r * n + i];
        if (zi)
            for (int64_t j = 0; j < k - 1; ++j) full[j] += (double)zi[r * (k - 1) + j];
        for (int64_t i = 0; i < n; ++i) y[r * n + i] = full[i];
        if (zf)
            for (int64_t j = 0; j < k - 1; ++j) zf[r * (k - 1) + j] = full[n + j];
    }
    free(full);
}

/* ------------------------------------------------------------------------- */
/* SciPy upfirdn (_upfirdn_apply.pyx).                                        */
/* ------------------------------------------------------------------------- */

/* _output_len (pyx:59-67): ceil(((in_len-1)*up + len_h)/down) in int64. */
ORACLE_API int64_t oracle_upfirdn_out_len(int64_t len_h, int64_t in_len, int64_t up,
                                          int64_t down)
{
    return (((in_len - 1) * up + len_h) - 1) / down + 1;
}

/* One row, mode='constant', cval=0, restating the state machine of _apply_impl
 * (pyx:421-481) on the transposed/flipped/padded taps of _pad_h (_upfirdn.py:47-64):
 * state (x_idx, t); per output: t += down; x_idx += t / up; t %= up; each output
 * accumulates oldest-sample-first over the h_per_phase taps of phase t, skipping
 * samples outside [0, len_x).  ACC is the accumulation type. */
#define DEFINE_UPFIRDN_ROW(NAME, TIN, ACC)                                              \
static void NAME(const TIN *h, int64_t len_h, const TIN *x, int64_t len_x, int64_t up,  \
                 int64_t down, ACC *out, int64_t len_out)                               \
{                                                                                       \
    const int64_t hpp = (len_h + up - 1) / up;                /* taps per phase */      \
    TIN *htf = (TIN *)calloc((size_t)(hpp * up), sizeof(TIN));                          \
    for (int64_t p = 0; p < up; ++p)                                                    \
        for (int64_t j = 0; j < hpp; ++j) {                                             \
            int64_t src = p + (hpp - 1 - j) * up;             /* flipped within phase */\
            htf[p * hpp + j] = (src < len_h) ? h[src] : (TIN)0;                         \
        }                                                                               \
    const int64_t padded_len = len_x + hpp - 1;                                         \
    int64_t x_idx = 0, t = 0, y_idx = 0;                                                \
    while (x_idx < padded_len && y_idx < len_out) {                                     \
        int64_t h_idx = t * hpp;                                                        \
        ACC acc = (ACC)0;                                                               \
        for (int64_t xc = x_idx - hpp + 1; xc <= x_idx; ++xc, ++h_idx) {                \
            if (xc < 0 || xc >= len_x) continue;              /* zero padding */        \
            ACC prod = (ACC)x[xc] * (ACC)htf[h_idx];                                    \
            acc = acc + prod;                                                           \
        }                                                                               \
        out[y_idx++] = acc;                                                             \
        t += down;                                                                      \
        x_idx += t / up;                                                                \
        t %= up;                                                                        \
    }                                                                                   \
    for (; y_idx < len_out; ++y_idx) out[y_idx] = (ACC)0;                               \
    free(htf);                                                                          \
}

DEFINE_UPFIRDN_ROW(upfirdn_row_f32, float, float)
DEFINE_UPFIRDN_ROW(upfirdn_row_f32_acc64, float, double)
DEFINE_UPFIRDN_ROW(upfirdn_row_f64, double, double)

ORACLE_API void oracle_upfirdn_f32(const float *h, int64_t len_h, const float *x,
                                   int64_t batch, int64_t len_x, int64_t up, int64_t down,
                                   float *y)
{
    const int64_t lo = oracle_upfirdn_out_len(len_h, len_x, up, down);
    for (int64_t r = 0; r < batch; ++r)
        upfirdn_row_f32(h, len_h, x + r * len_x, len_x, up, down, y + r * lo, lo);
}

ORACLE_API void oracle_upfirdn_f32_acc64(const float *h, int64_t len_h, const float *x,
                                         int64_t batch, int64_t len_x, int64_t up,
                                         int64_t down, double *y)
{
    const int64_t lo = oracle_upfirdn_out_len(len_h, len_x, up, down);
    for (int64_t r = 0; r < batch; ++r)
        upfirdn_row_f32_acc64(h, len_h, x + r * len_x, len_x, up, down, y + r * lo, lo);
}

ORACLE_API void oracle_upfirdn_f64(const double *h, int64_t len_h, const double *x,
                                   int64_t batch, int64_t len_x, int64_t up, int64_t down,
                                   double *y)
{
    const int64_t lo = oracle_upfirdn_out_len(len_h, len_x, up, down);
    for (int64_t r = 0; r < batch; ++r)
        upfirdn_row_f64(h, len_h, x + r * len_x, len_x, up, down, y + r * lo, lo);
}

/* ------------------------------------------------------------------------- */
/* SciPy upfirdn signal-extension modes (_upfirdn_apply.pyx:76-231, :421-481).  */
/* ------------------------------------------------------------------------- */
/* MODE enum of pyx:77-86 */
enum { OMODE_CONSTANT = 0, OMODE_SYMMETRIC = 1, OMODE_CONSTANT_EDGE = 2, OMODE_SMOOTH = 3, OMODE_PERIODIC = 4,
       OMODE_REFLECT = 5, OMODE_ANTISYMMETRIC = 6, OMODE_ANTIREFLECT = 7, OMODE_LINE = 8 };

/* C's % truncates like Cython's with cdivision(True); every left operand below is >= 0. */
#define DEFINE_EXTEND(SUF, T)                                                                        \
static T extend_left_##SUF(const T *x, int64_t idx, int64_t len_x, int mode, T cval)   /* pyx:110-170 */ \
{                                                                                                    \
    T le, lin_slope;                                                                                 \
    switch (mode) {                                                                                  \
    case OMODE_SYMMETRIC:                                                                            \
        if ((-idx) < len_x) return x[-idx - 1];                                                      \
        idx = (-idx - 1) % (2 * len_x);                                                              \
        return (idx < len_x) ? x[idx] : x[len_x - 1 - (idx - len_x)];                                \
    case OMODE_REFLECT:                                                                              \
        if ((-idx) < (len_x - 1)) return x[-idx];                                                    \
        idx = (-idx - 1) % (2 * (len_x - 1));                                                        \
        return (idx < (len_x - 1)) ? x[idx + 1] : x[len_x - 2 - (idx - (len_x - 1))];                \
    case OMODE_PERIODIC:                                                                             \
        idx = (-idx - 1) % len_x;                                                                    \
        return x[len_x - idx - 1];                                                                   \
    case OMODE_SMOOTH:                                                                               \
        return x[0] + (T)idx * (x[1] - x[0]);                                                        \
    case OMODE_LINE:                                                                                 \
        lin_slope = (x[len_x - 1] - x[0]) / (T)(len_x - 1);                                          \
        return x[0] + (T)idx * lin_slope;                                                            \
    case OMODE_ANTISYMMETRIC:                                                                        \
        if ((-idx) < len_x) return -x[-idx - 1];                                                     \
        idx = (-idx - 1) % (2 * len_x);                                                              \
        return (idx < len_x) ? -x[idx] : x[len_x - 1 - (idx - len_x)];                               \
    case OMODE_ANTIREFLECT:                                                                          \
        if ((-idx) < len_x) return x[0] - (x[-idx] - x[0]);                                          \
        le = x[0] + (x[0] - x[len_x - 1]) * (T)((-(idx) - 1) / (len_x - 1));                         \
        idx = (-idx - 1) % (2 * (len_x - 1));                                                        \
        return (idx < (len_x - 1)) ? le - (x[idx + 1] - x[0])                                        \
                                   : le - (x[len_x - 1] - x[len_x - 2 - (idx - (len_x - 1))]);       \
    case OMODE_CONSTANT_EDGE: return x[0];                                                           \
    case OMODE_CONSTANT: return cval;                                                                \
    default: return (T)-1;                                                                           \
    }                                                                                                \
}                                                                                                    \
static T extend_right_##SUF(const T *x, int64_t idx, int64_t len_x, int mode, T cval)  /* pyx:176-231 */ \
{                                                                                                    \
    T re, lin_slope;                                                                                 \
    switch (mode) {                                                                                  \
    case OMODE_SYMMETRIC:                                                                            \
        if (idx < (2 * len_x)) return x[len_x - 1 - (idx - len_x)];                                  \
        idx = idx % (2 * len_x);                                                                     \
        return (idx < len_x) ? x[idx] : x[len_x - 1 - (idx - len_x)];                                \
    case OMODE_REFLECT:                                                                              \
        if (idx < (2 * len_x - 1)) return x[len_x - 2 - (idx - len_x)];                              \
        idx = idx % (2 * (len_x - 1));                                                               \
        return (idx < (len_x - 1)) ? x[idx] : x[len_x - 1 - (idx - (len_x - 1))];                    \
    case OMODE_PERIODIC: return x[idx % len_x];                                                      \
    case OMODE_SMOOTH:                                                                               \
        return x[len_x - 1] + (T)(idx - len_x + 1) * (x[len_x - 1] - x[len_x - 2]);                  \
    case OMODE_LINE:                                                                                 \
        lin_slope = (x[len_x - 1] - x[0]) / (T)(len_x - 1);                                          \
        return x[len_x - 1] + (T)(idx - len_x + 1) * lin_slope;                                      \
    case OMODE_CONSTANT_EDGE: return x[len_x - 1];                                                   \
    case OMODE_ANTISYMMETRIC:                                                                        \
        if (idx < (2 * len_x)) return -x[len_x - 1 - (idx - len_x)];                                 \
        idx = idx % (2 * len_x);                                                                     \
        return (idx < len_x) ? x[idx] : -x[len_x - 1 - (idx - len_x)];                               \
    case OMODE_ANTIREFLECT:                                                                          \
        if (idx < (2 * len_x - 1)) return x[len_x - 1] - (x[len_x - 2 - (idx - len_x)] - x[len_x - 1]); \
        re = x[len_x - 1] + (x[len_x - 1] - x[0]) * (T)(idx / (len_x - 1) - 1);                      \
        idx = idx % (2 * (len_x - 1));                                                               \
        return (idx < (len_x - 1)) ? re + (x[idx] - x[0])                                            \
                                   : re + (x[len_x - 1] - x[len_x - 1 - (idx - (len_x - 1))]);       \
    case OMODE_CONSTANT: return cval;                                                                \
    default: return (T)-1;                                                                           \
    }                                                                                                \
}                                                                                                    \
/* _apply_impl (pyx:421-481) with a signal-extension mode: same state machine as above, samples      \
 * outside [0, len_x) come from extend_left / extend_right instead of being skipped. */              \
static void upfirdn_mode_row_##SUF(const T *h, int64_t len_h, const T *x, int64_t len_x, int64_t up, \
                                   int64_t down, int mode, T cval, double *out, int64_t len_out)     \
{                                                                                                    \
    const int64_t hpp = (len_h + up - 1) / up;                                                       \
    T *htf = (T *)calloc((size_t)(hpp * up), sizeof(T));                                             \
    for (int64_t p = 0; p < up; ++p)                                                                 \
        for (int64_t j = 0; j < hpp; ++j) {                                                          \
            int64_t src = p + (hpp - 1 - j) * up;                                                    \
            htf[p * hpp + j] = (src < len_h) ? h[src] : (T)0;                                        \
        }                                                                                            \
    const int64_t padded_len = len_x + hpp - 1;                                                      \
    int64_t x_idx = 0, t = 0, y_idx = 0;                                                             \
    while (x_idx < padded_len && y_idx < len_out) {                                                  \
        int64_t h_idx = t * hpp;                                                                     \
        double acc = 0.0;                                                                            \
        for (int64_t xc = x_idx - hpp + 1; xc <= x_idx; ++xc, ++h_idx) {                             \
            T xv;                                                                                    \
            if (xc < 0) xv = extend_left_##SUF(x, xc, len_x, mode, cval);                            \
            else if (xc >= len_x) xv = extend_right_##SUF(x, xc, len_x, mode, cval);                 \
            else xv = x[xc];                                                                         \
            acc += (double)xv * (double)htf[h_idx];                                                  \
        }                                                                                            \
        out[y_idx++] = acc;                                                                          \
        t += down;                                                                                   \
        x_idx += t / up;                                                                             \
        t %= up;                                                                                     \
    }                                                                                                \
    for (; y_idx < len_out; ++y_idx) out[y_idx] = 0.0;                                               \
    free(htf);                                                                                       \
}

DEFINE_EXTEND(f32, float)
DEFINE_EXTEND(f64, double)

/* f32 data (extension values formed in f32, as SciPy does for f32 input), products and sums in f64 */
ORACLE_API void oracle_upfirdn_mode_f32_acc64(const float *h, int64_t len_h, const float *x, int64_t batch,
                                              int64_t len_x, int64_t up, int64_t down, int mode, double cval,
                                              double *y)
{
    const int64_t lo = oracle_upfirdn_out_len(len_h, len_x, up, down);
    for (int64_t r = 0; r < batch; ++r)
        upfirdn_mode_row_f32(h, len_h, x + r * len_x, len_x, up, down, mode, (float)cval, y + r * lo, lo);
}

/* all-f64 twin: pins the restatement against SciPy's own f64 outputs to rounding */
ORACLE_API void oracle_upfirdn_mode_f64(const double *h, int64_t len_h, const double *x, int64_t batch,
                                        int64_t len_x, int64_t up, int64_t down, int mode, double cval,
                                        double *y)
{
    const int64_t lo = oracle_upfirdn_out_len(len_h, len_x, up, down);
    for (int64_t r = 0; r < batch; ++r)
        upfirdn_mode_row_f64(h, len_h, x + r * len_x, len_x, up, down, mode, cval, y + r * lo, lo);
}

/* ------------------------------------------------------------------------- */
/* SciPy resample_poly with an array `window` (_signaltools.py:3865-3957).    */
/* The plan is pure int64 arithmetic and must be bit-exact.                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int64_t up, down;          /* after gcd reduction (:3882-3884) */
    int64_t n_out;             /* ceil(n_in*up/down)  (:3887-3889) */
    int64_t half_len;          /* (len_h-1)/2         (:3895)      */
    int64_t n_pre_pad;         /* down - half_len%down (:3912)     */
    int64_t n_post_pad;        /* grown by the while loop (:3916)  */
    int64_t n_pre_remove;      /* (half_len+n_pre_pad)/down (:3914)*/
    int64_t len_h_padded;      /* len_h + n_pre_pad + n_post_pad   */
    int64_t upfirdn_len;       /* _output_len of the padded filter */
} oracle_resample_plan;

static int64_t gcd64(int64_t a, int64_t b)
{
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}

ORACLE_API void oracle_resample_poly_plan(int64_t n_in, int64_t len_h, int64_t up,
                                          int64_t down, oracle_resample_plan *p)
{
    int64_t g = gcd64(up, down);
    up /= g;
    down /= g;
    p->up = up;
    p->down = down;
    int64_t n_out = n_in * up;
    p->n_out = n_out / down + (n_out % down != 0);
    p->half_len = (len_h - 1) / 2;
    p->n_pre_pad = down - p->half_len % down;
    p->n_post_pad = 0;
    p->n_pre_remove = (p->half_len + p->n_pre_pad) / down;
    while (oracle_upfirdn_out_len(len_h + p->n_pre_pad + p->n_post_pad, n_in, up, down) <
           p->n_out + p->n_pre_remove)
        p->n_post_pad += 1;
    p->len_h_padded = len_h + p->n_pre_pad + p->n_post_pad;
    p->upfirdn_len = oracle_upfirdn_out_len(p->len_h_padded, n_in, up, down);
}

/* y gets plan.n_out values per row.  `window` is the user filter; SciPy scales it by
 * `up` in the data dtype (f32 here, :3909) before padding.  acc64 != 0 accumulates in
 * f64 (judge), else f32 like SciPy's float32 kernel. */
ORACLE_API void oracle_resample_poly_f32(const float *window, int64_t len_h, int64_t up,
                                         int64_t down, const float *x, int64_t batch,
                                         int64_t n_in, int acc64, double *y)
{
    oracle_resample_plan p;
    oracle_resample_poly_plan(n_in, len_h, up, down, &p);
    if (p.up == 1 && p.down == 1) {                      /* :3885-3886 copy */
        for (int64_t i = 0; i < batch * n_in; ++i) y[i] = x[i];
        return;
    }
    float *h = (float *)calloc((size_t)p.len_h_padded, sizeof(float));
    for (int64_t i = 0; i < len_h; ++i) h[p.n_pre_pad + i] = window[i] * (float)p.up;
    double *full64 = (double *)malloc((size_t)p.upfirdn_len * sizeof(double));
    float *full32 = (float *)malloc((size_t)p.upfirdn_len * sizeof(float));
    for (int64_t r = 0; r < batch; ++r) {
        if (acc64) {
            upfirdn_row_f32_acc64(h, p.len_h_padded, x + r * n_in, n_in, p.up, p.down,
                                  full64, p.upfirdn_len);
            for (int64_t i = 0; i < p.n_out; ++i) y[r * p.n_out + i] = full64[p.n_pre_remove + i];
        } else {
            upfirdn_row_f32(h, p.len_h_padded, x + r * n_in, n_in, p.up, p.down, full32,
                            p.upfirdn_len);
            for (int64_t i = 0; i < p.n_out; ++i) y[r * p.n_out + i] = full32[p.n_pre_remove + i];
        }
    }
    free(h);
    free(full64);
    free(full32);
}

/* ------------------------------------------------------------------------- */
/* SciPy filtfilt(b, [1], x), method='pad' (_signaltools.py:4745-4790).       */
/* padtype: 0 none, 1 odd, 2 even, 3 constant.  padlen<0 -> 3*ntaps (:4804).   */
/* Returns 0, or -1 when len(x) <= edge (:4809 raises ValueError).            */
/* zi = lfilter_zi(b,[1]) has the closed form zi[j] = sum_{m>j} b[m]           */
/* (solution of (I-A^T) zi = b[1:], :4309-4313, for a=[1]).                   */
/* All arithmetic f64 on f32 inputs.                                          */
/* ------------------------------------------------------------------------- */
static void lfilter_fir_zi_f64(const double *b, int64_t k, const double *x, int64_t n,
                               const double *zi, double x0scale, double *y)
{
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        int64_t dmax = (i < k - 1) ? i : (k - 1);
        for (int64_t d = 0; d <= dmax; ++d) acc += b[d] * x[i - d];
        if (i < k - 1) acc += zi[i] * x0scale;
        y[i] = acc;
    }
}

ORACLE_API int oracle_filtfilt_fir_f32(const float *b, int64_t k, int padtype, int64_t padlen,
                                       const float *x, int64_t batch, int64_t n, double *y)
{
    int64_t edge = (padtype == 0) ? 0 : (padlen < 0 ? 3 * k : padlen);
    if (n <= edge) return -1;
    const int64_t ne = n + 2 * edge;
    double *bd = (double *)malloc((size_t)k * sizeof(double));
    double *zi = (double *)calloc((size_t)(k > 1 ? k - 1 : 1), sizeof(double));
    double *ext = (double *)malloc((size_t)ne * sizeof(double));
    double *y1 = (double *)malloc((size_t)ne * sizeof(double));
    double *rev = (double *)malloc((size_t)ne * sizeof(double));
    for (int64_t i = 0; i < k; ++i) bd[i] = b[i];
    for (int64_t j = k - 2; j >= 0; --j) zi[j] = (j + 1 < k - 1 ? zi[j + 1] : 0.0) + bd[j + 1];
    for (int64_t r = 0; r < batch; ++r) {
        const float *xr = x + r * n;
        for (int64_t i = 0; i < n; ++i) ext[edge + i] = xr[i];
        for (int64_t j = 0; j < edge; ++j) {
            /* left: samples x[edge], x[edge-1], ..., x[1]; right: x[n-2], ..., x[n-edge-1] */
            double l = xr[edge - j], rr = xr[n - 2 - j];
            if (padtype == 1) {          /* odd_ext (_arraytools.py:57-107) */
                ext[j] = 2.0 * (double)xr[0] - l;
                ext[edge + n + j] = 2.0 * (double)xr[n - 1] - rr;
            } else if (padtype == 2) {   /* even_ext */
                ext[j] = l;
                ext[edge + n + j] = rr;
            } else {                     /* const_ext */
                ext[j] = xr[0];
                ext[edge + n + j] = xr[n - 1];
            }
        }
        lfilter_fir_zi_f64(bd, k, ext, ne, zi, ext[0], y1);
        for (int64_t i = 0; i < ne; ++i) rev[i] = y1[ne - 1 - i];
        lfilter_fir_zi_f64(bd, k, rev, ne, zi, rev[0], y1);
        for (int64_t i = 0; i < n; ++i) y[r * n + i] = y1[ne - 1 - (edge + i)];
    }
    free(bd);
    free(zi);
    free(ext);
    free(y1);
    free(rev);
    return 0;
}

ORACLE_API const char *oracle_version(void) { return "scir-b200 oracle 0.1 (test infrastructure)"; }
