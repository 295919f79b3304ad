// ext_modes.cuh -- SciPy's signal-extension modes for upfirdn / resample_poly (scipy/signal/
// _upfirdn_apply.pyx:76-231): the value of the virtual sample x[idx] for idx outside [0, n), for f32 and f64 rows.
// Only edge tiles call this (interior tiles arrive by bulk copy), so clarity beats speed; arithmetic is in the
// dtype of x without FMA contraction, in SciPy's order of operations, because SciPy forms these values in the
// dtype of x.
#pragma once
#include "../../include/scir_b200.h"

namespace scir_b200 {

template <typename T>
struct ExtSpecT {
    int mode;        // SCIR_B200_EXT_*
    T cval;          // EXT_CONSTANT only
};
using ExtSpec = ExtSpecT<float>;
using ExtSpec64 = ExtSpecT<double>;

__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__device__ __forceinline__ T upfirdn_ext_left(const T* __restrict__ x, long long idx, long long n, const ExtSpecT<T>& e)
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
        case SCIR_B200_EXT_SMOOTH: return rn_add(x[0], rn_mul(static_cast<T>(idx), rn_sub(x[1], x[0])));
        case SCIR_B200_EXT_LINE: {
            const T slope = rn_div(rn_sub(x[n - 1], x[0]), static_cast<T>(n - 1));
            return rn_add(x[0], rn_mul(static_cast<T>(idx), slope));
        }
        case SCIR_B200_EXT_ANTISYMMETRIC: {
            if (-idx < n) return -x[-idx - 1];
            const long long j = (-idx - 1) % (2 * n);
            return (j < n) ? -x[j] : x[n - 1 - (j - n)];
        }
        case SCIR_B200_EXT_ANTIREFLECT: {
            if (-idx < n) return rn_sub(x[0], rn_sub(x[-idx], x[0]));
            const T le = rn_add(x[0], rn_mul(rn_sub(x[0], x[n - 1]), static_cast<T>((-idx - 1) / (n - 1))));
            const long long j = (-idx - 1) % (2 * (n - 1));
            return (j < n - 1) ? rn_sub(le, rn_sub(x[j + 1], x[0])) : rn_sub(le, rn_sub(x[n - 1], x[n - 2 - (j - (n - 1))]));
        }
        case SCIR_B200_EXT_EDGE: return x[0];
        default: return e.cval;
    }
}

template <typename T>
__device__ __forceinline__ T upfirdn_ext_right(const T* __restrict__ x, long long idx, long long n, const ExtSpecT<T>& e)
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
            return rn_add(x[n - 1], rn_mul(static_cast<T>(idx - n + 1), rn_sub(x[n - 1], x[n - 2])));
        case SCIR_B200_EXT_LINE: {
            const T slope = rn_div(rn_sub(x[n - 1], x[0]), static_cast<T>(n - 1));
            return rn_add(x[n - 1], rn_mul(static_cast<T>(idx - n + 1), slope));
        }
        case SCIR_B200_EXT_EDGE: return x[n - 1];
        case SCIR_B200_EXT_ANTISYMMETRIC: {
            if (idx < 2 * n) return -x[n - 1 - (idx - n)];
            const long long j = idx % (2 * n);
            return (j < n) ? x[j] : -x[n - 1 - (j - n)];
        }
        case SCIR_B200_EXT_ANTIREFLECT: {
            if (idx < 2 * n - 1) return rn_sub(x[n - 1], rn_sub(x[n - 2 - (idx - n)], x[n - 1]));
            const T re = rn_add(x[n - 1], rn_mul(rn_sub(x[n - 1], x[0]), static_cast<T>(idx / (n - 1) - 1)));
            const long long j = idx % (2 * (n - 1));
            return (j < n - 1) ? rn_add(re, rn_sub(x[j], x[0])) : rn_add(re, rn_sub(x[n - 1], x[n - 1 - (j - (n - 1))]));
        }
        default: return e.cval;
    }
}

// virtual sample at any index
template <typename T>
__device__ __forceinline__ T upfirdn_sample(const T* __restrict__ x, long long idx, long long n, const ExtSpecT<T>& e)
{
    if (idx < 0) return upfirdn_ext_left(x, idx, n, e);
    if (idx >= n) return upfirdn_ext_right(x, idx, n, e);
    return x[idx];
}

}  // namespace scir_b200
