// upfirdn_poly.cu -- tiled polyphase kernel (placeholder until the tile kernel lands: declines).
#include "common.cuh"

namespace scir_b200 {

int launch_upfirdn_poly(scir_b200_ctx*, const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t,
                        int64_t, float*, int64_t, int64_t, int64_t, bool* handled)
{
    *handled = false;
    return SCIR_B200_OK;
}

}  // namespace scir_b200
