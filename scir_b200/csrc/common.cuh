// common.cuh -- internal declarations shared by the translation units of libscir_b200.so.
// Not part of the ABI (that is include/scir_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>

#include "../../include/scir_b200.h"

namespace scir_b200 {

// ---- errors -----------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);          // records thread-local message, returns code
int cuda_error(cudaError_t e, const char* what);        // maps a CUDA error to an ABI code
void clear_error();

#define SCIR_CUDA(call, what)                                              \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) return ::scir_b200::cuda_error(_e, what);   \
    } while (0)

#define SCIR_TRY(expr)                       \
    do {                                     \
        int _rc = (expr);                    \
        if (_rc != SCIR_B200_OK) return _rc; \
    } while (0)

// Every ABI entry point that takes a ctx opens with SCIR_ENTER(ctx): NULL check, then the ctx's device is made
// current for the duration of the call and the CALLER's current context is put back on return (DeviceScope), so a
// call on a cuda:1 ctx inside a PyTorch process never moves torch's current device.
#define SCIR_ENTER(ctx)                                    \
    SCIR_TRY(::scir_b200::check_ctx(ctx));                 \
    ::scir_b200::DeviceScope _scir_scope((ctx)->device);   \
    SCIR_TRY(_scir_scope.rc)

// Saves the calling thread's current CUDA context, binds `device`, restores on destruction.  The driver's
// cuCtxGetCurrent / cuCtxSetCurrent (resolved through cudaGetDriverEntryPoint: no link-time libcuda dependency)
// are used rather than cudaGetDevice / cudaSetDevice, because restoring "device 0" with cudaSetDevice in a thread
// that never touched the GPU would CREATE a primary context on device 0 (CUDA >= 12 initialises eagerly).
struct DeviceScope {
    void* prev = nullptr;      // CUcontext
    bool restore = false;
    int rc = SCIR_B200_OK;
    explicit DeviceScope(int device);
    ~DeviceScope();
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};

// ---- context ----------------------------------------------------------------------------------
struct Options {
    int64_t variant = 0;        // 0 auto; 1 force generic (non-bulk) tile IO; 2 naive 1-thread/output
    int64_t host_block_rows = 0; // rows per block in the *_host streaming paths (0 = auto)
    int64_t long_tap_path = 0;  // 0 auto (see launch_fir); 1 force FP32 direct; 2 force tcgen05 Toeplitz; 3 force overlap-save FFT
    int64_t os_packed = 0;      // overlap-save, N = 16384: 1 packed-lane kernel (two butterflies per thread in FADD2/FMUL2/FFMA2: 31 % fewer
                                // instructions, same 5.9 ms on config 3 -- the kernel is latency / barrier bound at one CTA per SM, not
                                // issue bound; profiles/README.md), 0 scalar kernel (default)
    int64_t os_min_k = 512;     // auto mode: from this tap count on the overlap-save FFT path is COSTED against the tensor kernel
                                // (api.cu: launch_fir) and taken when cheaper.  Measured (profiles/README.md): K = 509 6.5 ms tensor
                                // vs 9.2 ms FFT (config 5), K = 4097 17.0 ms tensor vs 5.8 ms FFT (config 3); the models meet at
                                // K ~ 760 for long rows, later for rows shorter than a block pair
    int64_t toeplitz_terms = 3; // split products per tap: 3 (hh,hm,mh), 4 (+mm), 6 (+hl,lh)
    int64_t toeplitz_split = 0; // operand format of the split: 0 block-scaled FP16 (11-bit terms), 1 BF16 (8-bit terms)
    int64_t toeplitz_chains = 1; // accumulation chains per tile in TMEM (2: consecutive MMAs alternate accumulators; measured: no gain)
    int64_t filtfilt_fused = 1;  // padded filtfilt as ONE zero-phase pass with b (*) flip(b) when the pad covers k-1 samples (0: two passes)
    int64_t ffma2 = 1;           // FP32 direct kernels, K <= 256: packed FFMA2 core (0: scalar FFMA, the A/B arm)
    int64_t toeplitz_tn = 0;     // tile width of the Toeplitz kernel: 0 auto, 64 or 128 columns
    int64_t toeplitz_tn_short = 128; // auto: width for filters with K <= 129 (64 = deeper prefetch; measured in profiles/README.md)
    int64_t toeplitz_stcs = 0;   // 1: evict-first hint on the epilogue's output stores (A/B)
    int64_t toeplitz_adaptive_budget = 500; // tensor kernel: thousandths of the tolerance 1e-5 * sum|c| * max|x| that dropping the two mid-term
                                 // products of small-tap Toeplitz blocks may cost in the worst case (fir_toeplitz.cu: make_plan); 0 = never drop
    int64_t toeplitz_ts = 1;     // 1: keep the first Toeplitz blocks in TMEM (A operand from TMEM); 0: all operands from shared memory
    int64_t toeplitz_loader = 0; // 0 auto (TMA-fed in-place buffers when they fit); 1 force the register-prefetch loader
    int64_t toeplitz_min_k = 1024; // auto mode: tap counts from here on always take the tensor path
    int64_t toeplitz_min_k_full = 2;  // auto mode: ... and from here on when the cost model (api.cu: prefer_toeplitz) says so
    int64_t upfirdn_variant = 0; // 0 auto; 1 force generic polyphase kernel (others: upfirdn_poly.cu launch_one)
    int64_t upfirdn_ws_stages = 3; // input stages of the warp-specialised polyphase kernel: 3 (2 CTAs/SM) or 2 (3 CTAs/SM)
    // *_host entry points, when the caller's arrays are PAGEABLE (a Rust Vec / ndarray / numpy buffer):
    int64_t host_stage = 1;      // 1 stage through the ctx's pinned ring with copy threads (default); 0 hand the pageable
                                 // pointer to cudaMemcpyAsync (driver-staged, host-synchronous: the A/B arm);
                                 // 2 cudaHostRegister the caller's spans for the duration of the call
    int64_t host_copy_threads = 0;   // copy threads of the pinned ring (0 = auto: min(8, hardware threads / 2))
    int64_t host_stage_wc = 0;   // 1: the H2D half of the pinned ring is write-combined memory
    int64_t host_stage_nt = 1;   // 1: the copy threads use non-temporal stores (no read-for-ownership); 0: plain memcpy (A/B)
};

class CopyPool;                 // copy_pool.hpp

struct DeviceBuffer {           // RAII-free growable scratch owned by the ctx
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace scir_b200

struct scir_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 0;
    int max_smem_optin = 0;
    uint64_t launches = 0;
    uint64_t toeplitz_launches = 0;        // launches served by the tcgen05 Toeplitz kernel
    int toeplitz_last_mma_per_tile = 0;    // MMAs per tile of the last Toeplitz launch, and how many of its blocks ran hi x hi only
    int toeplitz_last_hh_blocks = 0;
    uint64_t filtfilt_fused_calls = 0;     // filtfilt calls served by the single-pass form
    uint64_t fixup_launches = 0;           // non-finite fix-up kernels (one after every FIR launch; idle on finite data)
    uint64_t poly_launches = 0;            // launches served by the polyphase TILE kernel (tests)
    scir_b200::Options opt;
    scir_b200::DeviceBuffer scratch;       // filtfilt intermediate etc.
    scir_b200::DeviceBuffer toep_flags;    // per-tile non-finite flags of the last Toeplitz launch
    scir_b200::DeviceBuffer toep_taps;     // taps of the Toeplitz kernel (device copy) ...
    std::vector<float> toep_taps_host;     // ... and what it currently holds
    scir_b200::DeviceBuffer gen_taps;      // phase-transposed taps of the any-rate tiled polyphase kernel ...
    std::vector<float> gen_taps_host;      // ... and what the buffer currently holds
    uint64_t gen_tiled_launches = 0;       // launches served by it
    scir_b200::DeviceBuffer row_bg;        // resample_poly padtype statistics: one float per row
    // overlap-save FFT path (fir_os.cu): twiddles, the current filter's taps and spectrum, and what they were made from
    scir_b200::DeviceBuffer os_tw, os_taps, os_H;
    std::vector<float> os_taps_host;
    int os_logn = 0;
    uint64_t os_launches = 0;              // launches served by the overlap-save kernel
    // *_host streaming pipeline resources (lazily created): a ring of kHostSlots row blocks
    static constexpr int kHostSlots = 6;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    scir_b200::DeviceBuffer stage_in[kHostSlots], stage_out[kHostSlots];
    cudaEvent_t ev_in[kHostSlots] = {};
    cudaEvent_t ev_k[kHostSlots] = {};
    cudaEvent_t ev_out[kHostSlots] = {};
    // pageable callers: pinned twins of the ring and the threads that fill / drain them
    scir_b200::DeviceBuffer pin_in[kHostSlots], pin_out[kHostSlots];
    bool pin_in_wc = false;
    scir_b200::CopyPool* pool = nullptr;
    uint64_t host_staged_calls = 0;        // *_host calls that went through the pinned ring
    uint64_t host_registered_calls = 0;    // ... that registered the caller's spans instead
};

namespace scir_b200 {

int ctx_bind(const scir_b200_ctx* ctx);                                  // cudaSetDevice
int ctx_scratch(scir_b200_ctx* ctx, DeviceBuffer& buf, size_t bytes);    // grow-only
int check_ctx(const scir_b200_ctx* ctx);

// ---- the FIR pass: one launch of the direct-form tile kernel -----------------------------------
// Virtual input sequence v[i], i in [0, n_v): v[i] = ext(x_row)[i + in_off]; outside [0, n_v) the
// sequence is zero (BOUND_ZERO) or held at its end value (BOUND_HOLD).
//   dir=+1 (causal):     out[i] = sum_d c[d] * v[i - d]
//   dir=-1 (anticausal): out[i] = sum_d c[d] * v[i + d]
// for i in [out_begin, out_end), written to y_row[i + out_off].
enum { EXT_NONE = 0, EXT_ODD = 1, EXT_EVEN = 2, EXT_CONST = 3 };
enum { BOUND_ZERO = 0, BOUND_HOLD = 1 };

struct FirPass {
    const float* x;
    float* y;
    long long ld_x, ld_y;
    long long batch;
    long long n_x;            // valid underlying indices are [0, n_x) (used by ext modes)
    long long n_v;            // virtual length
    long long in_off, out_off;
    long long out_begin, out_end;
    int ext_mode;
    int bound;
    int dir;
};

struct FirPass64 {            // the same pass on f64 rows (fir_f64.cu)
    const double* x;
    double* y;
    long long ld_x, ld_y;
    long long batch;
    long long n_x, n_v;
    long long in_off, out_off;
    long long out_begin, out_end;
    int ext_mode;
    int bound;
    int dir;
};

// c: coefficients by delay index (lfilter order), length k.
// launch_fir picks the kernel (tcgen05 Toeplitz vs FP32 direct) from the ctx options; launch_fir_pass
// is the FP32 direct family.
int launch_fir(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k);
int launch_fir_pass(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k);

// lfilter streaming-state helpers (small kernels)
int launch_add_zi(scir_b200_ctx* ctx, float* d_y, int64_t ld_y, const float* d_zi, int64_t batch,
                  int64_t n, int64_t k);
int launch_compute_zf(scir_b200_ctx* ctx, const float* b, int64_t k, const float* d_x, int64_t ld_x,
                      const float* d_zi, float* d_zf, int64_t batch, int64_t n);

// ---- polyphase upfirdn ------------------------------------------------------------------------
// ext_mode: SCIR_B200_EXT_* (samples outside [0, n_in)), cval for EXT_CONSTANT
int launch_upfirdn(scir_b200_ctx* ctx, const float* h, int64_t len_h, int64_t up, int64_t down,
                   const float* d_x, int64_t ld_x, int64_t batch, int64_t n_in, float* d_y,
                   int64_t ld_y, int64_t m_begin, int64_t m_count, int ext_mode = 0, float cval = 0.f);

// ---- long-tap tensor-core path (tcgen05 block-Toeplitz) -----------------------------------------
bool toeplitz_supported(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k, int64_t* tiles = nullptr, bool* aligned = nullptr);
int launch_fir_toeplitz(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k);

// ---- long-tap overlap-save path (FP32 shared-memory FFT, fir_os.cu) -------------------------------------------
bool fir_os_supported(const scir_b200_ctx* ctx, const FirPass& pass, int64_t k, double* est_seconds = nullptr);
int launch_fir_os(scir_b200_ctx* ctx, const FirPass& pass, const float* c, int64_t k);

int64_t upfirdn_out_len(int64_t len_h, int64_t in_len, int64_t up, int64_t down);

// ---- f64 twin of the hot path (fir_f64.cu); taps by delay index, host pointer -----------------------------
int launch_fir_f64(scir_b200_ctx* ctx, const double* d_x, int64_t ld_x, const double* taps_by_delay, int64_t k, double* d_y,
                   int64_t ld_y, int64_t batch, int64_t n);
int launch_fir_pass_f64(scir_b200_ctx* ctx, const FirPass64& pass, const double* taps_by_delay, int64_t k);

// ---- DeviceArray elementwise ops (elementwise.cu): op 0 add-scalar, 1 mul-scalar, 2 add ------------------
int launch_elementwise(scir_b200_ctx* ctx, int op, const float* d_a, const float* d_b, float alpha, float* d_y, int64_t n);
// per-row mean (0) / minimum (1) / maximum (2), and y = x + sign * bg[row]
int launch_row_stat(scir_b200_ctx* ctx, int stat, const float* d_x, int64_t ld_x, int64_t batch, int64_t n, float* d_out);
int launch_row_offset(scir_b200_ctx* ctx, const float* d_x, int64_t ld_x, const float* d_bg, float sign, float* d_y,
                      int64_t ld_y, int64_t batch, int64_t n);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace scir_b200
