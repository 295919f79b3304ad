// mg.cu -- in-process multi-GPU front end: contiguous row blocks per device, one ctx (stream) and
// one host thread per device, no collective (rows are independent: crates/scir-gpu/src/lib.rs:1138-1140
// has no cross-row state).  The one-process-per-GPU launcher in scir_b200/dist.py shards with the same
// scir_b200_shard_rows().
#include "common.cuh"

#include <thread>

struct scir_b200_mg {
    std::vector<scir_b200_ctx*> ctxs;
};

using namespace scir_b200;

// Runs fn(ctx, first_row, rows) for every shard on its own host thread; the first failure wins.
template <typename Fn>
static int mg_for_each_shard(scir_b200_mg* mg, int64_t batch, Fn&& fn)
{
    if (!mg) return set_error(SCIR_B200_ERR_INVALID_ARG, "mg is NULL");
    const int world = static_cast<int>(mg->ctxs.size());
    std::vector<int> rcs(world, SCIR_B200_OK);
    std::vector<std::string> msgs(world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r) {
        threads.emplace_back([&, r]() {
            int64_t r0 = 0, r1 = 0;
            rcs[r] = scir_b200_shard_rows(batch, world, r, &r0, &r1);
            if (rcs[r] == SCIR_B200_OK && r1 > r0) rcs[r] = fn(mg->ctxs[r], r0, r1 - r0);
            if (rcs[r] != SCIR_B200_OK) msgs[r] = scir_b200_last_error();   // thread-local: carry it out
        });
    }
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
        if (rcs[r] != SCIR_B200_OK) return set_error(rcs[r], "shard %d: %s", r, msgs[r].c_str());
    return SCIR_B200_OK;
}

extern "C" {

int scir_b200_mg_create(const int* devices, int n_devices, scir_b200_mg** mg)
{
    if (!mg) return set_error(SCIR_B200_ERR_INVALID_ARG, "mg out-pointer is NULL");
    *mg = nullptr;
    if (!devices || n_devices < 1) return set_error(SCIR_B200_ERR_INVALID_ARG, "need at least one device");
    scir_b200_mg* m = new scir_b200_mg();
    for (int i = 0; i < n_devices; ++i) {
        scir_b200_ctx* c = nullptr;
        int rc = scir_b200_ctx_create(devices[i], &c);
        if (rc != SCIR_B200_OK) {
            for (auto* p : m->ctxs) scir_b200_ctx_destroy(p);
            delete m;
            return rc;
        }
        m->ctxs.push_back(c);
    }
    *mg = m;
    return SCIR_B200_OK;
}

int scir_b200_mg_destroy(scir_b200_mg* mg)
{
    if (!mg) return SCIR_B200_OK;
    for (auto* p : mg->ctxs) scir_b200_ctx_destroy(p);
    delete mg;
    return SCIR_B200_OK;
}

int scir_b200_mg_device_count(const scir_b200_mg* mg, int* n_devices)
{
    if (!mg || !n_devices) return set_error(SCIR_B200_ERR_INVALID_ARG, "NULL argument");
    *n_devices = static_cast<int>(mg->ctxs.size());
    return SCIR_B200_OK;
}

int scir_b200_mg_fir1d_batched_f32_host(scir_b200_mg* mg, const float* h_x, int64_t ld_x, const float* taps,
                                        int64_t k, int tap_order, float* h_y, int64_t ld_y, int64_t batch,
                                        int64_t n)
{
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int64_t r0, int64_t rows) {
        return scir_b200_fir1d_batched_f32_host(ctx, h_x + r0 * ld_x, ld_x, taps, k, tap_order, h_y + r0 * ld_y, ld_y, rows, n);
    });
}

int scir_b200_mg_resample_poly_f32_host(scir_b200_mg* mg, const float* window, int64_t len_h, int64_t up, int64_t down,
                                        const float* h_x, int64_t ld_x, int64_t batch, int64_t n_in, float* h_y,
                                        int64_t ld_y)
{
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int64_t r0, int64_t rows) {
        return scir_b200_resample_poly_f32_host(ctx, window, len_h, up, down, h_x + r0 * ld_x, ld_x, rows, n_in,
                                                h_y + r0 * ld_y, ld_y);
    });
}

int scir_b200_mg_filtfilt_fir_f32_host(scir_b200_mg* mg, const float* b, int64_t k, int pad_mode, int64_t padlen,
                                       const float* h_x, int64_t ld_x, float* h_y, int64_t ld_y, int64_t batch,
                                       int64_t n)
{
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int64_t r0, int64_t rows) {
        return scir_b200_filtfilt_fir_f32_host(ctx, b, k, pad_mode, padlen, h_x + r0 * ld_x, ld_x, h_y + r0 * ld_y, ld_y,
                                               rows, n);
    });
}

}  // extern "C"
