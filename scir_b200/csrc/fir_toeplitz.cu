// fir_toeplitz.cu -- long-tap tcgen05 block-Toeplitz path (not built yet: declines every request,
// so long taps run on the FP32 direct kernel).
#include "common.cuh"

namespace scir_b200 {

bool toeplitz_supported(const scir_b200_ctx*, const FirPass&, int64_t) { return false; }

int launch_fir_toeplitz(scir_b200_ctx*, const FirPass&, const float*, int64_t)
{
    return set_error(SCIR_B200_ERR_UNSUPPORTED, "tcgen05 Toeplitz path not built");
}

}  // namespace scir_b200
