// mg.cu -- in-process multi-GPU front end: contiguous row blocks per device, one ctx (stream) and
// one PERSISTENT host thread per device, no collective on the data path (rows are independent:
// crates/scir-gpu/src/lib.rs:1138-1140 has no cross-row state).  The one-process-per-GPU launcher in
// scir_b200/dist.py shards with the same scir_b200_shard_rows().
//
// The optional "whole output on one device" step (north_star (d)) is scir_b200_mg_gather_rows_f32:
// a peer-to-peer fan-in over NVLink, one cudaMemcpy2DAsync per shard on the SOURCE device's stream (so it is ordered
// after that shard's kernel), never part of the filtering calls.
#include "common.cuh"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

using namespace scir_b200;

namespace {

// One long-lived thread per device: the launch code keeps per-thread state (opt-in shared-memory sizes, staged tap
// blocks), which a fresh std::thread per call would rebuild -- and leak -- every time.
class ShardWorker {
public:
    ShardWorker() : th_([this]() { loop(); }) {}
    ~ShardWorker()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> f)
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            task_ = std::move(f);
            busy_ = true;
        }
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this]() { return !busy_; });
    }

private:
    void loop()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this]() { return stop_ || (busy_ && task_); });
                if (stop_) return;
                f = std::move(task_);
                task_ = nullptr;
            }
            f();
            {
                std::lock_guard<std::mutex> lk(m_);
                busy_ = false;
            }
            cv_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::function<void()> task_;
    bool busy_ = false, stop_ = false;
    std::thread th_;
};

}  // namespace

struct scir_b200_mg {
    std::vector<scir_b200_ctx*> ctxs;
    std::vector<std::unique_ptr<ShardWorker>> workers;
    std::mutex call_mutex;                          // one mg call at a time (the workers hold one task each)
};

// Runs fn(ctx, shard, first_row, rows) for every shard on that device's thread; the first failure wins.
template <typename Fn>
static int mg_for_each_shard(scir_b200_mg* mg, int64_t batch, Fn&& fn)
{
    if (!mg) return set_error(SCIR_B200_ERR_INVALID_ARG, "mg is NULL");
    std::lock_guard<std::mutex> call(mg->call_mutex);
    const int world = static_cast<int>(mg->ctxs.size());
    std::vector<int> rcs(world, SCIR_B200_OK);
    std::vector<std::string> msgs(world);
    for (int r = 0; r < world; ++r) {
        mg->workers[r]->post([&, r]() {
            int64_t r0 = 0, r1 = 0;
            rcs[r] = scir_b200_shard_rows(batch, world, r, &r0, &r1);
            if (rcs[r] == SCIR_B200_OK && r1 > r0) rcs[r] = fn(mg->ctxs[r], r, r0, r1 - r0);
            if (rcs[r] != SCIR_B200_OK) msgs[r] = scir_b200_last_error();   // thread-local: carry it out
        });
    }
    for (int r = 0; r < world; ++r) mg->workers[r]->wait();
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
    const int hw = std::max(2, static_cast<int>(std::thread::hardware_concurrency()));
    for (int i = 0; i < n_devices; ++i) {
        scir_b200_ctx* c = nullptr;
        int rc = scir_b200_ctx_create(devices[i], &c);
        if (rc != SCIR_B200_OK) {
            for (auto* p : m->ctxs) scir_b200_ctx_destroy(p);
            delete m;
            return rc;
        }
        // pageable callers: every device has its own copy threads; share the host's cores between them
        c->opt.host_copy_threads = std::max(2, std::min(8, hw / (2 * n_devices)));
        m->ctxs.push_back(c);
    }
    // peer access for the gather fan-in (NVLink); failure only means the copies are staged by the driver
    for (int i = 0; i < n_devices; ++i) {
        DeviceScope scope(devices[i]);
        for (int j = 0; j < n_devices; ++j) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess) cudaGetLastError();          // already enabled is fine
            } else {
                cudaGetLastError();
            }
        }
    }
    for (int i = 0; i < n_devices; ++i) m->workers.emplace_back(new ShardWorker());
    *mg = m;
    return SCIR_B200_OK;
}

int scir_b200_mg_destroy(scir_b200_mg* mg)
{
    if (!mg) return SCIR_B200_OK;
    mg->workers.clear();
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

int scir_b200_mg_ctx(const scir_b200_mg* mg, int shard, scir_b200_ctx** ctx)
{
    if (!mg || !ctx) return set_error(SCIR_B200_ERR_INVALID_ARG, "NULL argument");
    if (shard < 0 || shard >= static_cast<int>(mg->ctxs.size()))
        return set_error(SCIR_B200_ERR_INVALID_ARG, "shard %d out of range (have %zu)", shard, mg->ctxs.size());
    *ctx = mg->ctxs[static_cast<size_t>(shard)];
    return SCIR_B200_OK;
}

int scir_b200_mg_sync(scir_b200_mg* mg)
{
    if (!mg) return set_error(SCIR_B200_ERR_INVALID_ARG, "mg is NULL");
    for (auto* c : mg->ctxs) SCIR_TRY(scir_b200_ctx_sync(c));
    return SCIR_B200_OK;
}

int scir_b200_mg_fir1d_batched_f32_host(scir_b200_mg* mg, const float* h_x, int64_t ld_x, const float* taps,
                                        int64_t k, int tap_order, float* h_y, int64_t ld_y, int64_t batch,
                                        int64_t n)
{
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int, int64_t r0, int64_t rows) {
        return scir_b200_fir1d_batched_f32_host(ctx, h_x + r0 * ld_x, ld_x, taps, k, tap_order, h_y + r0 * ld_y, ld_y, rows, n);
    });
}

int scir_b200_mg_resample_poly_f32_host(scir_b200_mg* mg, const float* window, int64_t len_h, int64_t up, int64_t down,
                                        const float* h_x, int64_t ld_x, int64_t batch, int64_t n_in, float* h_y,
                                        int64_t ld_y)
{
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int, int64_t r0, int64_t rows) {
        return scir_b200_resample_poly_f32_host(ctx, window, len_h, up, down, h_x + r0 * ld_x, ld_x, rows, n_in,
                                                h_y + r0 * ld_y, ld_y);
    });
}

int scir_b200_mg_filtfilt_fir_f32_host(scir_b200_mg* mg, const float* b, int64_t k, int pad_mode, int64_t padlen,
                                       const float* h_x, int64_t ld_x, float* h_y, int64_t ld_y, int64_t batch,
                                       int64_t n)
{
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int, int64_t r0, int64_t rows) {
        return scir_b200_filtfilt_fir_f32_host(ctx, b, k, pad_mode, padlen, h_x + r0 * ld_x, ld_x, h_y + r0 * ld_y, ld_y,
                                               rows, n);
    });
}

// Device-resident shards: d_x[s] / d_y[s] point at shard s's first row ON DEVICE s (scir_b200_shard_rows(batch, world, s)
// rows of pitch ld_x[s] / ld_y[s]).  Asynchronous on every device's stream; scir_b200_mg_sync() waits for all of them.
int scir_b200_mg_fir1d_batched_f32(scir_b200_mg* mg, const float* const* d_x, const int64_t* ld_x, const float* taps,
                                   int64_t k, int tap_order, float* const* d_y, const int64_t* ld_y, int64_t batch,
                                   int64_t n)
{
    if (!d_x || !d_y || !ld_x || !ld_y) return set_error(SCIR_B200_ERR_INVALID_ARG, "NULL shard table");
    return mg_for_each_shard(mg, batch, [&](scir_b200_ctx* ctx, int s, int64_t, int64_t rows) {
        return scir_b200_fir1d_batched_f32(ctx, d_x[s], ld_x[s], taps, k, tap_order, d_y[s], ld_y[s], rows, n);
    });
}

// Fan the row shards in to ONE device: d_dst (batch, n) of pitch ld_dst on the device of shard `dst_shard`.
// Each shard's copy is queued on its own device's stream behind that shard's kernel and travels peer-to-peer
// (NVLink / NVSwitch when peer access could be enabled); the call returns when every copy has landed.
int scir_b200_mg_gather_rows_f32(scir_b200_mg* mg, const float* const* d_shards, const int64_t* ld_shards, int dst_shard,
                                 float* d_dst, int64_t ld_dst, int64_t batch, int64_t n)
{
    if (!mg) return set_error(SCIR_B200_ERR_INVALID_ARG, "mg is NULL");
    if (!d_shards || !ld_shards) return set_error(SCIR_B200_ERR_INVALID_ARG, "NULL shard table");
    const int world = static_cast<int>(mg->ctxs.size());
    if (dst_shard < 0 || dst_shard >= world) return set_error(SCIR_B200_ERR_INVALID_ARG, "dst_shard %d out of range", dst_shard);
    if (batch < 0 || n < 0) return set_error(SCIR_B200_ERR_INVALID_ARG, "negative shape");
    if (batch > 0 && n > 0 && !d_dst) return set_error(SCIR_B200_ERR_INVALID_ARG, "d_dst is NULL");
    if (batch > 1 && ld_dst < n) return set_error(SCIR_B200_ERR_INVALID_ARG, "ld_dst < n");
    if (batch == 0 || n == 0) return SCIR_B200_OK;
    for (int s = 0; s < world; ++s) {
        int64_t r0 = 0, r1 = 0;
        SCIR_TRY(scir_b200_shard_rows(batch, world, s, &r0, &r1));
        if (r1 == r0) continue;
        if (!d_shards[s]) return set_error(SCIR_B200_ERR_INVALID_ARG, "shard %d pointer is NULL", s);
        scir_b200_ctx* c = mg->ctxs[static_cast<size_t>(s)];
        DeviceScope scope(c->device);
        SCIR_TRY(scope.rc);
        SCIR_CUDA(cudaMemcpy2DAsync(d_dst + r0 * ld_dst, static_cast<size_t>(ld_dst) * 4, d_shards[s],
                                    static_cast<size_t>(ld_shards[s]) * 4, static_cast<size_t>(n) * 4,
                                    static_cast<size_t>(r1 - r0), cudaMemcpyDefault, c->stream),
                  "cudaMemcpy2DAsync(gather fan-in)");
    }
    return scir_b200_mg_sync(mg);
}

}  // extern "C"
