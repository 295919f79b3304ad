// scir_b200.hpp -- C++ host-side mirror of the reference's Rust surface for the batched-FIR path,
// header-only, over the C ABI in scir_b200.h.
//
// The reference is compiled code (Rust) and its toolchain is not in this image, so this is the
// host side a C++ caller uses; a Rust maintainer binds the same C symbols.  Names,
// argument meaning and error behaviour follow the reference:
//
//   scir::gpu::DType, Device, GpuError, DeviceArray<T>      crates/scir-gpu/src/lib.rs:16-35, 57-84, 95-190
//   scir::gpu::fir1d_batched_f32_cuda(x, taps)              lib.rs:1036-1113   (Result -> throws GpuError)
//   scir::gpu::fir1d_batched_f32_auto(x, taps, device)      lib.rs:515-531     (no silent CPU fallback)
//   scir::signal::gpu::fir1d_batched_f32(x, taps, device)   crates/scir-signal/src/lib.rs:365-375
//   scir::signal::{lfilter, upfirdn, resample_poly, filtfilt, filtfilt_zero_state}
//                                                           SciPy semantics (scipy/signal/_signaltools.py,
//                                                           _upfirdn.py) -- the FIR routes north_star names
//
// Arrays are row-major (batch, n) float matrices like ndarray::Array2<f32> in standard layout.
#ifndef SCIR_B200_HPP
#define SCIR_B200_HPP

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "scir_b200.h"

namespace scir {
namespace gpu {

enum class DType { F32, F64 };          // lib.rs:16-22
enum class Device { Cpu, Cuda };        // lib.rs:25-35 (no wgpu arm: north_star forbids multi-backend dispatch)

// lib.rs:57-74
class GpuError : public std::runtime_error {
public:
    enum class Kind { BackendUnavailable, ShapeMismatch };
    GpuError(Kind kind, const std::string& msg, int code)
        : std::runtime_error(kind == Kind::BackendUnavailable ? "backend not available: " + msg
                                                              : "shape mismatch" + (msg.empty() ? "" : ": " + msg)),
          kind_(kind), code_(code) {}
    Kind kind() const { return kind_; }
    int code() const { return code_; }

private:
    Kind kind_;
    int code_;
};

inline void check(int rc)
{
    if (rc == SCIR_B200_OK) return;
    const std::string msg = scir_b200_last_error();
    if (rc == SCIR_B200_ERR_SHAPE) throw GpuError(GpuError::Kind::ShapeMismatch, msg, rc);
    if (rc == SCIR_B200_ERR_INVALID_ARG) throw std::invalid_argument(msg);      // the reference asserts (lib.rs:841-842)
    throw GpuError(GpuError::Kind::BackendUnavailable, msg, rc);
}

// Row-major owning matrix (stand-in for ndarray::Array2<f32>).
struct Array2 {
    std::size_t rows = 0, cols = 0;
    std::vector<float> data;
    Array2() = default;
    Array2(std::size_t r, std::size_t c, float fill = 0.f) : rows(r), cols(c), data(r * c, fill) {}
    Array2(std::size_t r, std::size_t c, std::vector<float> v) : rows(r), cols(c), data(std::move(v))
    {
        if (data.size() != r * c) throw GpuError(GpuError::Kind::ShapeMismatch, "data length != rows*cols", SCIR_B200_ERR_SHAPE);
    }
    float& operator()(std::size_t r, std::size_t c) { return data[r * cols + c]; }
    float operator()(std::size_t r, std::size_t c) const { return data[r * cols + c]; }
    int64_t ld() const { return static_cast<int64_t>(cols ? cols : 1); }
};

// A long-lived handle (device + stream + scratch); replaces the per-call CudaCtx, lib.rs:601-622.
class Context {
public:
    explicit Context(int device = 0) { check(scir_b200_ctx_create(device, &ctx_)); }
    ~Context() { scir_b200_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    scir_b200_ctx* get() const { return ctx_; }
    void sync() const { check(scir_b200_ctx_sync(ctx_)); }
    void set_option(const char* key, int64_t v) { check(scir_b200_ctx_set_option(ctx_, key, v)); }

private:
    scir_b200_ctx* ctx_ = nullptr;
};

inline Context& default_context()
{
    thread_local std::unique_ptr<Context> c;
    if (!c) {                              // the thread's CURRENT device; throws GpuError::BackendUnavailable without a B200
        int dev = 0;
        check(scir_b200_current_device(&dev));
        c.reset(new Context(dev));
    }
    return *c;
}

inline int device_count()
{
    int n = 0;
    check(scir_b200_device_count(&n));
    return n;
}

// lib.rs:1036-1113.  y[b,i] = sum_t taps[k-1-t] * x[b,i-t], same shape as x.
inline Array2 fir1d_batched_f32_cuda(const Array2& x, const std::vector<float>& taps)
{
    Array2 y(x.rows, x.cols);
    check(scir_b200_fir1d_batched_f32_host(default_context().get(), x.data.data(), x.ld(), taps.data(),
                                           static_cast<int64_t>(taps.size()), SCIR_B200_TAPS_SCIR, y.data.data(), y.ld(),
                                           static_cast<int64_t>(x.rows), static_cast<int64_t>(x.cols)));
    return y;
}

// The crate's CPU function (lib.rs:1134-1152): f32 multiply then f32 add, newest sample first.  Reached only when
// the caller names Device::Cpu; nothing on the Device::Cuda arm can end up here.
inline Array2 fir1d_batched_f32(const Array2& x, const std::vector<float>& taps)
{
    Array2 y(x.rows, x.cols);
    const std::size_t k = taps.size();
    for (std::size_t b = 0; b < x.rows; ++b)
        for (std::size_t i = 0; i < x.cols; ++i) {
            volatile float acc = 0.f;                       // volatile: no contraction of the product into the sum
            for (std::size_t t = 0; t < k && t <= i; ++t) {
                volatile float prod = taps[k - 1 - t] * x.data[b * x.cols + i - t];
                acc = acc + prod;
            }
            y.data[b * x.cols + i] = acc;
        }
    return y;
}

// lib.rs:515-531 minus the silent fallback (lib.rs:520-523): Device::Cuda runs on the B200 or throws;
// Device::Cpu is the crate's own CPU function, as in the reference (lib.rs:517-518).
inline Array2 fir1d_batched_f32_auto(const Array2& x, const std::vector<float>& taps, Device device)
{
    if (device == Device::Cuda) return fir1d_batched_f32_cuda(x, taps);
    return fir1d_batched_f32(x, taps);
}

// lib.rs:77-190 with real device storage (f32 on the device).
class DeviceArray {
public:
    static DeviceArray from_cpu_slice(std::vector<std::size_t> shape, DType dtype, const std::vector<float>& data)
    {
        std::size_t n = 1;
        for (auto s : shape) n *= s;
        if (n != data.size()) throw std::invalid_argument("shape product != data length");     // assert_eq!, lib.rs:96
        DeviceArray a;
        a.shape_ = std::move(shape);
        a.dtype_ = dtype;
        a.host_ = data;
        return a;
    }
    DeviceArray() = default;
    DeviceArray(DeviceArray&& o) noexcept { *this = std::move(o); }
    DeviceArray& operator=(DeviceArray&& o) noexcept
    {
        release();
        shape_ = std::move(o.shape_); dtype_ = o.dtype_; device_ = o.device_; host_ = std::move(o.host_);
        dptr_ = o.dptr_; o.dptr_ = nullptr; o.device_ = Device::Cpu;
        return *this;
    }
    ~DeviceArray() { release(); }
    const std::vector<std::size_t>& shape() const { return shape_; }
    DType dtype() const { return dtype_; }
    Device device() const { return device_; }
    std::size_t len() const { std::size_t n = 1; for (auto s : shape_) n *= s; return n; }
    std::vector<float> to_cpu_vec() const
    {
        if (device_ == Device::Cpu) return host_;
        std::vector<float> out(len());
        check(scir_b200_memcpy_d2h(default_context().get(), out.data(), dptr_, out.size() * sizeof(float)));
        return out;
    }
    void to_device(Device device)
    {
        if (device == device_) return;
        if (device == Device::Cuda) {
            if (dtype_ != DType::F32) throw GpuError(GpuError::Kind::BackendUnavailable, "only f32 arrays live on the device", SCIR_B200_ERR_UNSUPPORTED);
            void* p = nullptr;
            check(scir_b200_malloc(default_context().get(), std::max<std::size_t>(host_.size(), 1) * sizeof(float), &p));
            check(scir_b200_memcpy_h2d(default_context().get(), p, host_.data(), host_.size() * sizeof(float)));
            dptr_ = p;
            device_ = Device::Cuda;
        } else {
            host_ = to_cpu_vec();
            release();
        }
    }
    // Elementwise ops with device dispatch (lib.rs:268-377).  Device::Cuda arrays run the kernels on the resident
    // pointers; Device::Cpu arrays run the crate's own loops (lib.rs:206-255) -- the array's device decides, there
    // is no fallback from one to the other.  Both give the same bits (one add or mul, no FMA).
    DeviceArray add_scalar_auto(float alpha) const
    {
        if (device_ == Device::Cpu) return host_map([alpha](float a) { return a + alpha; });
        DeviceArray out = like("add_scalar_auto");
        check(scir_b200_add_scalar_f32(default_context().get(), static_cast<const float*>(dptr_), alpha,
                                       static_cast<float*>(out.dptr_), static_cast<int64_t>(len())));
        return out;
    }
    DeviceArray mul_scalar_auto(float alpha) const
    {
        if (device_ == Device::Cpu) return host_map([alpha](float a) { return a * alpha; });
        DeviceArray out = like("mul_scalar_auto");
        check(scir_b200_mul_scalar_f32(default_context().get(), static_cast<const float*>(dptr_), alpha,
                                       static_cast<float*>(out.dptr_), static_cast<int64_t>(len())));
        return out;
    }
    DeviceArray add_auto(const DeviceArray& other) const
    {
        if (shape_ != other.shape_) throw GpuError(GpuError::Kind::ShapeMismatch, "add_auto: shapes differ", SCIR_B200_ERR_SHAPE);   // lib.rs:304-306
        if (other.device_ != device_)
            throw GpuError(GpuError::Kind::BackendUnavailable, "add_auto: both arrays must be on the same device", SCIR_B200_ERR_NO_DEVICE);
        if (device_ == Device::Cpu) {
            DeviceArray out = host_map([](float a) { return a; });
            for (std::size_t i = 0; i < out.host_.size(); ++i) out.host_[i] = host_[i] + other.host_[i];
            return out;
        }
        DeviceArray out = like("add_auto");
        check(scir_b200_add_f32(default_context().get(), static_cast<const float*>(dptr_), static_cast<const float*>(other.dptr_),
                                static_cast<float*>(out.dptr_), static_cast<int64_t>(len())));
        return out;
    }
    // Device-resident FIR: no PCIe traffic when chaining.
    DeviceArray fir1d_batched(const std::vector<float>& taps, int tap_order = SCIR_B200_TAPS_SCIR) const
    {
        if (device_ != Device::Cuda || shape_.size() != 2) throw GpuError(GpuError::Kind::ShapeMismatch, "need a 2-D array on Device::Cuda", SCIR_B200_ERR_SHAPE);
        DeviceArray out;
        out.shape_ = shape_; out.dtype_ = dtype_;
        const int64_t b = static_cast<int64_t>(shape_[0]), n = static_cast<int64_t>(shape_[1]);
        void* p = nullptr;
        check(scir_b200_malloc(default_context().get(), std::max<std::size_t>(len(), 1) * sizeof(float), &p));
        out.dptr_ = p; out.device_ = Device::Cuda;
        check(scir_b200_fir1d_batched_f32(default_context().get(), static_cast<const float*>(dptr_), n ? n : 1, taps.data(),
                                          static_cast<int64_t>(taps.size()), tap_order, static_cast<float*>(p), n ? n : 1, b, n));
        return out;
    }

private:
    template <typename F>
    DeviceArray host_map(F f) const                         // Device::Cpu arm: the crate's own loop
    {
        DeviceArray out;
        out.shape_ = shape_; out.dtype_ = dtype_; out.host_.resize(host_.size());
        for (std::size_t i = 0; i < host_.size(); ++i) out.host_[i] = f(host_[i]);
        return out;
    }
    DeviceArray like(const char* what) const               // same shape, fresh device storage
    {
        if (device_ != Device::Cuda)
            throw GpuError(GpuError::Kind::BackendUnavailable,
                           std::string(what) + ": Device::Cpu arrays are served by the reference crate's CPU loops (lib.rs:206-255)",
                           SCIR_B200_ERR_NO_DEVICE);
        DeviceArray out;
        out.shape_ = shape_; out.dtype_ = dtype_;
        void* p = nullptr;
        check(scir_b200_malloc(default_context().get(), std::max<std::size_t>(len(), 1) * sizeof(float), &p));
        out.dptr_ = p; out.device_ = Device::Cuda;
        return out;
    }
    void release()
    {
        if (dptr_) scir_b200_free(default_context().get(), dptr_);
        dptr_ = nullptr;
        device_ = Device::Cpu;
    }
    std::vector<std::size_t> shape_;
    DType dtype_ = DType::F32;
    Device device_ = Device::Cpu;
    std::vector<float> host_;
    void* dptr_ = nullptr;
};

}  // namespace gpu

namespace signal {

namespace gpu {
// crates/scir-signal/src/lib.rs:372-374
inline scir::gpu::Array2 fir1d_batched_f32(const scir::gpu::Array2& x, const std::vector<float>& taps, scir::gpu::Device device)
{
    return scir::gpu::fir1d_batched_f32_auto(x, taps, device);
}
}  // namespace gpu

using scir::gpu::Array2;
using scir::gpu::check;
using scir::gpu::default_context;

// lfilter(b, [a0], x): scipy/signal/_signaltools.py:2181-2242 (FIR branch)
inline Array2 lfilter(const std::vector<float>& b, float a0, const Array2& x)
{
    if (a0 == 0.f) throw std::invalid_argument("a[0] must be nonzero");
    std::vector<float> bs(b);
    for (auto& v : bs) v /= a0;
    Array2 y(x.rows, x.cols);
    check(scir_b200_fir1d_batched_f32_host(default_context().get(), x.data.data(), x.ld(), bs.data(), static_cast<int64_t>(bs.size()),
                                           SCIR_B200_TAPS_LFILTER, y.data.data(), y.ld(), static_cast<int64_t>(x.rows),
                                           static_cast<int64_t>(x.cols)));
    return y;
}

inline int64_t upfirdn_output_len(int64_t len_h, int64_t in_len, int64_t up, int64_t down)
{
    return scir_b200_upfirdn_out_len(len_h, in_len, up, down);          // _upfirdn_apply.pyx:59-67
}

inline scir_b200_resample_plan resample_poly_plan(int64_t n_in, int64_t len_h, int64_t up, int64_t down)
{
    scir_b200_resample_plan p;
    check(scir_b200_resample_poly_plan(n_in, len_h, up, down, &p));     // _signaltools.py:3882-3918
    return p;
}

// resample_poly(x, up, down, window=h), padtype='constant': _signaltools.py:3865-3957
inline Array2 resample_poly(const Array2& x, int64_t up, int64_t down, const std::vector<float>& window)
{
    const auto plan = resample_poly_plan(static_cast<int64_t>(x.cols), static_cast<int64_t>(window.size()), up, down);
    const std::size_t n_out = (plan.up == 1 && plan.down == 1) ? x.cols : static_cast<std::size_t>(plan.n_out);
    Array2 y(x.rows, n_out);
    check(scir_b200_resample_poly_f32_host(default_context().get(), window.data(), static_cast<int64_t>(window.size()), up, down,
                                           x.data.data(), x.ld(), static_cast<int64_t>(x.rows), static_cast<int64_t>(x.cols),
                                           y.data.data(), y.ld()));
    return y;
}

enum class Pad { ZeroState = SCIR_B200_PAD_ZERO_STATE, Odd = SCIR_B200_PAD_ODD, Even = SCIR_B200_PAD_EVEN,
                 Constant = SCIR_B200_PAD_CONSTANT, None = SCIR_B200_PAD_SCIPY_NONE };

// filtfilt(b, [1], x, padtype, padlen): _signaltools.py:4745-4826; Pad::ZeroState is the reference's own
// unpadded forward-backward structure (crates/scir-signal/src/lib.rs:278-291) with an FIR numerator.
inline Array2 filtfilt(const std::vector<float>& b, const Array2& x, Pad pad = Pad::Odd, int64_t padlen = -1)
{
    Array2 y(x.rows, x.cols);
    const int rc = scir_b200_filtfilt_fir_f32_host(default_context().get(), b.data(), static_cast<int64_t>(b.size()),
                                                   static_cast<int>(pad), padlen, x.data.data(), x.ld(), y.data.data(), y.ld(),
                                                   static_cast<int64_t>(x.rows), static_cast<int64_t>(x.cols));
    if (rc == SCIR_B200_ERR_SHAPE) throw std::invalid_argument(scir_b200_last_error());     // SciPy raises ValueError, :4809
    check(rc);
    return y;
}

inline Array2 filtfilt_zero_state(const std::vector<float>& b, const Array2& x) { return filtfilt(b, x, Pad::ZeroState); }

}  // namespace signal
}  // namespace scir

#endif  // SCIR_B200_HPP
