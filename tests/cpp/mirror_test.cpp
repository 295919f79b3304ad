// Exercises include/scir_b200.hpp (the C++ mirror of the reference's Rust surface) end to end.
//   mirror_test cpu : no GPU expected -- Device::Cuda must fail loudly (GpuError::BackendUnavailable),
//                     integer plans must still work (they are pure host code).
//   mirror_test gpu : the reference's known-answer vector (crates/scir-gpu/src/lib.rs:1251-1260) and the
//                     scir-signal routes through the host-array entry points.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "scir_b200.hpp"

using namespace scir;

static int fails = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
            ++fails;                                                        \
        }                                                                   \
    } while (0)

static void integer_plans()
{
    // scipy/signal/tests/test_upfirdn.py:311-322
    EXPECT(signal::upfirdn_output_len(1001, 100000000LL, 320, 441) == 72562360LL);
    const auto p = signal::resample_poly_plan(1 << 20, 96, 3, 2);      // BASELINE config 4
    EXPECT(p.n_out == 1572864 && p.half_len == 47 && p.n_pre_pad == 1 && p.n_post_pad == 0 && p.n_pre_remove == 24);
    EXPECT(p.upfirdn_len == 1572911);
    int64_t r0 = 0, r1 = 0;
    EXPECT(scir_b200_shard_rows(10, 4, 1, &r0, &r1) == 0 && r0 == 3 && r1 == 6);
}

int main(int argc, char** argv)
{
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    integer_plans();
    const gpu::Array2 x(2, 4, {1.f, 2.f, 3.f, 4.f, 0.5f, 0.f, -0.5f, -1.f});
    const std::vector<float> taps = {0.25f, 0.5f, 0.25f};
    if (mode == "cpu") {
        bool threw = false;
        try {
            (void)gpu::fir1d_batched_f32_auto(x, taps, gpu::Device::Cuda);
        } catch (const gpu::GpuError& e) {
            threw = e.kind() == gpu::GpuError::Kind::BackendUnavailable;
            std::printf("Device::Cuda without a GPU -> %s\n", e.what());
        }
        EXPECT(threw);                                                   // never a silent CPU fallback (lib.rs:520-523)
        // Device::Cpu, asked for explicitly, is the crate's own CPU function (lib.rs:517-518): the reference's
        // known-answer vector (lib.rs:1251-1260)
        const gpu::Array2 yc = gpu::fir1d_batched_f32_auto(x, taps, gpu::Device::Cpu);
        const float want_cpu[8] = {0.25f, 1.f, 2.f, 3.f, 0.125f, 0.25f, 0.f, -0.5f};
        for (int i = 0; i < 8; ++i) EXPECT(std::fabs(yc.data[i] - want_cpu[i]) <= 1e-7f);
    } else {
        const gpu::Array2 y = signal::gpu::fir1d_batched_f32(x, taps, gpu::Device::Cuda);
        const float want[8] = {0.25f, 1.f, 2.f, 3.f, 0.125f, 0.25f, 0.f, -0.5f};
        for (int i = 0; i < 8; ++i) EXPECT(std::fabs(y.data[i] - want[i]) <= 1e-7f);
        // lfilter([1,1],[1],arange(6)) = [0,1,3,5,7,9]   (scipy test_signaltools.py:1848-1853)
        const gpu::Array2 a(1, 6, {0.f, 1.f, 2.f, 3.f, 4.f, 5.f});
        const gpu::Array2 l = signal::lfilter({1.f, 1.f}, 1.f, a);
        const float lw[6] = {0.f, 1.f, 3.f, 5.f, 7.f, 9.f};
        for (int i = 0; i < 6; ++i) EXPECT(l.data[i] == lw[i]);
        // filtfilt with the identity filter returns x (test_signaltools.py:2797-2804)
        gpu::Array2 s(1, 50);
        for (int i = 0; i < 50; ++i) s.data[i] = std::sin(0.3f * i);
        const gpu::Array2 f = signal::filtfilt({1.f}, s);
        for (int i = 0; i < 50; ++i) EXPECT(std::fabs(f.data[i] - s.data[i]) <= 1e-6f);
        bool threw = false;
        try {
            (void)signal::filtfilt(std::vector<float>(30, 0.1f), s);    // padlen 90 >= n: SciPy raises ValueError
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        EXPECT(threw);
        const gpu::Array2 r = signal::resample_poly(s, 2, 1, {0.25f, 0.5f, 0.25f});
        EXPECT(r.cols == 100);
        // DeviceArray: device-resident chain (lib.rs:77-190 with real device storage)
        gpu::DeviceArray d = gpu::DeviceArray::from_cpu_slice({2, 4}, gpu::DType::F32, x.data);
        d.to_device(gpu::Device::Cuda);
        EXPECT(d.device() == gpu::Device::Cuda);
        gpu::DeviceArray e = d.fir1d_batched(taps);
        const auto back = e.to_cpu_vec();
        for (int i = 0; i < 8; ++i) EXPECT(std::fabs(back[i] - want[i]) <= 1e-7f);
        // elementwise _auto ops: the reference's doc-test vectors (lib.rs:258-262, :293-298, :345-350), chained on the device
        gpu::DeviceArray da = gpu::DeviceArray::from_cpu_slice({3}, gpu::DType::F32, {1.0f, 2.0f, 3.0f});
        gpu::DeviceArray db = gpu::DeviceArray::from_cpu_slice({3}, gpu::DType::F32, {0.5f, 1.5f, 2.5f});
        // a Device::Cpu array runs the crate's own loop (lib.rs:206-221) and stays on the CPU; mixed devices are an error
        const gpu::DeviceArray dcpu = da.add_scalar_auto(1.0f);
        EXPECT(dcpu.device() == gpu::Device::Cpu && dcpu.to_cpu_vec() == (std::vector<float>{2.0f, 3.0f, 4.0f}));
        da.to_device(gpu::Device::Cuda);
        bool mixed_threw = false;
        try { (void)da.add_auto(db); } catch (const gpu::GpuError&) { mixed_threw = true; }
        EXPECT(mixed_threw);
        db.to_device(gpu::Device::Cuda);
        EXPECT(da.add_scalar_auto(1.0f).to_cpu_vec() == (std::vector<float>{2.0f, 3.0f, 4.0f}));
        EXPECT(da.mul_scalar_auto(2.0f).to_cpu_vec() == (std::vector<float>{2.0f, 4.0f, 6.0f}));
        EXPECT(da.add_auto(db).to_cpu_vec() == (std::vector<float>{1.5f, 3.5f, 5.5f}));
        EXPECT(da.mul_scalar_auto(2.0f).add_scalar_auto(-1.0f).add_auto(db).to_cpu_vec() == (std::vector<float>{1.5f, 4.5f, 7.5f}));
        gpu::DeviceArray dc = gpu::DeviceArray::from_cpu_slice({2}, gpu::DType::F32, {1.0f, 2.0f});
        dc.to_device(gpu::Device::Cuda);
        bool shape_threw = false;
        try { (void)da.add_auto(dc); } catch (const gpu::GpuError& e) { shape_threw = (e.kind() == gpu::GpuError::Kind::ShapeMismatch); }
        EXPECT(shape_threw);
    }
    std::printf("%s: %d failure(s)\n", mode.c_str(), fails);
    return fails ? 1 : 0;
}
