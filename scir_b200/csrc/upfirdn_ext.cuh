// upfirdn_ext.cuh -- SciPy's signal-extension modes for upfirdn / resample_poly (scipy/signal/
// _upfirdn_apply.pyx:76-231): the value of the virtual sample x[idx] for idx outside [0, n).  Only edge tiles
// call this (interior tiles arrive by bulk copy), so clarity beats speed; arithmetic is f32 without FMA
// contraction, in SciPy's order of operations, because SciPy forms these values in the dtype of x.
#pragma once
#include "../../include/scir_b200.h"

namespace scir_b200 {

struct ExtSpec {
    int mode;        // SCIR_B200_EXT_*
    float cval;      // EXT_CONSTANT only
};

__device__ __forceinline__ float upfirdn_ext_left(const float* __restrict__ x, long long idx, long long n, const ExtSpec& e)
{
    switch (e.mode) {                                                  // idx < 0
        case SCIR_B200_EXT_SYMMETRIC: {
            if (-idx < n) return x[-idx - 1];
            const long long j = (-idx - 1) % (2 * n);
            return (j < n) ? x[j] : x[n - 1 - (j - n)];
        }
        case SCIR_B200_EXT_REFLECT: {
            if (-idx < n - 1) return x[-idx];
            const long long j = (-idx - 1) % (2 * (n - 1));
            return (j < n - 1) ? x[j + 1] : x[n - 2 - (j - (n - 1))];
        }
        case SCIR_B200_EXT_PERIODIC: return x[n - ((-idx - 1) % n) - 1];
        case SCIR_B200_EXT_SMOOTH: return __fadd_rn(x[0], __fmul_rn(static_cast<float>(idx), __fsub_rn(x[1], x[0])));
        case SCIR_B200_EXT_LINE: {
            const float slope = __fdiv_rn(__fsub_rn(x[n - 1], x[0]), static_cast<float>(n - 1));
            return __fadd_rn(x[0], __fmul_rn(static_cast<float>(idx), slope));
        }
        case SCIR_B200_EXT_ANTISYMMETRIC: {
            if (-idx < n) return -x[-idx - 1];
            const long long j = (-idx - 1) % (2 * n);
            return (j < n) ? -x[j] : x[n - 1 - (j - n)];
        }
        case SCIR_B200_EXT_ANTIREFLECT: {
            if (-idx < n) return __fsub_rn(x[0], __fsub_rn(x[-idx], x[0]));
            const float le = __fadd_rn(x[0], __fmul_rn(__fsub_rn(x[0], x[n - 1]), static_cast<float>((-idx - 1) / (n - 1))));
            const long long j = (-idx - 1) % (2 * (n - 1));
            return (j < n - 1) ? __fsub_rn(le, __fsub_rn(x[j + 1], x[0])) : __fsub_rn(le, __fsub_rn(x[n - 1], x[n - 2 - (j - (n - 1))]));
        }
        case SCIR_B200_EXT_EDGE: return x[0];
        default: return e.cval;
    }
}

__device__ __forceinline__ float upfirdn_ext_right(const float* __restrict__ x, long long idx, long long n, const ExtSpec& e)
{
    switch (e.mode) {                                                  // idx >= n
        case SCIR_B200_EXT_SYMMETRIC: {
            if (idx < 2 * n) return x[n - 1 - (idx - n)];
            const long long j = idx % (2 * n);
            return (j < n) ? x[j] : x[n - 1 - (j - n)];
        }
        case SCIR_B200_EXT_REFLECT: {
            if (idx < 2 * n - 1) return x[n - 2 - (idx - n)];
            const long long j = idx % (2 * (n - 1));
            return (j < n - 1) ? x[j] : x[n - 1 - (j - (n - 1))];
        }
        case SCIR_B200_EXT_PERIODIC: return x[idx % n];
        case SCIR_B200_EXT_SMOOTH:
            return __fadd_rn(x[n - 1], __fmul_rn(static_cast<float>(idx - n + 1), __fsub_rn(x[n - 1], x[n - 2])));
        case SCIR_B200_EXT_LINE: {
            const float slope = __fdiv_rn(__fsub_rn(x[n - 1], x[0]), static_cast<float>(n - 1));
            return __fadd_rn(x[n - 1], __fmul_rn(static_cast<float>(idx - n + 1), slope));
        }
        case SCIR_B200_EXT_EDGE: return x[n - 1];
        case SCIR_B200_EXT_ANTISYMMETRIC: {
            if (idx < 2 * n) return -x[n - 1 - (idx - n)];
            const long long j = idx % (2 * n);
            return (j < n) ? x[j] : -x[n - 1 - (j - n)];
        }
        case SCIR_B200_EXT_ANTIREFLECT: {
            if (idx < 2 * n - 1) return __fsub_rn(x[n - 1], __fsub_rn(x[n - 2 - (idx - n)], x[n - 1]));
            const float re = __fadd_rn(x[n - 1], __fmul_rn(__fsub_rn(x[n - 1], x[0]), static_cast<float>(idx / (n - 1) - 1)));
            const long long j = idx % (2 * (n - 1));
            return (j < n - 1) ? __fadd_rn(re, __fsub_rn(x[j], x[0])) : __fadd_rn(re, __fsub_rn(x[n - 1], x[n - 1 - (j - (n - 1))]));
        }
        default: return e.cval;
    }
}

// virtual sample at any index
__device__ __forceinline__ float upfirdn_sample(const float* __restrict__ x, long long idx, long long n, const ExtSpec& e)
{
    if (idx < 0) return upfirdn_ext_left(x, idx, n, e);
    if (idx >= n) return upfirdn_ext_right(x, idx, n, e);
    return x[idx];
}

}  // namespace scir_b200
