// copy_pool.hpp -- a small pool of host threads that move bytes between caller memory and the ctx's pinned
// staging ring (api.cu: host_pipeline, pageable-memory route).  Jobs are plain memcpy ranges; a submission
// returns a ticket the issuing thread can wait on, and the waiter helps draining the queue instead of idling.
#pragma once

#include <emmintrin.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace scir_b200 {

// memcpy with non-temporal (streaming) stores.  A block on its way through the pinned ring is written once and next
// read by the DMA engine (or, on the way out, by nobody soon): ordinary stores would first READ every destination
// line into the cache (read-for-ownership) and evict useful data -- three bus transfers per byte copied instead of
// two.  On the B200 boxes' hosts the staged path is memory-bandwidth bound, so this is worth ~1.3x (profiles/README.md).
// SSE2 only (baseline x86-64); the write-combining buffers merge the 16-byte stores into full lines.
inline void stream_copy(char* dst, const char* src, size_t bytes)
{
    if (bytes < 4096) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
    if (head) {
        memcpy(dst, src, head);
        dst += head;
        src += head;
        bytes -= head;
    }
    const size_t lines = bytes / 64;
    __m128i* d = reinterpret_cast<__m128i*>(dst);
    const __m128i* s = reinterpret_cast<const __m128i*>(src);
    for (size_t i = 0; i < lines; ++i) {
        const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1);
        const __m128i c = _mm_loadu_si128(s + 4 * i + 2), e = _mm_loadu_si128(s + 4 * i + 3);
        _mm_stream_si128(d + 4 * i, a);
        _mm_stream_si128(d + 4 * i + 1, b);
        _mm_stream_si128(d + 4 * i + 2, c);
        _mm_stream_si128(d + 4 * i + 3, e);
    }
    _mm_sfence();
    const size_t done = lines * 64;
    if (bytes > done) memcpy(dst + done, src + done, bytes - done);
}

class CopyPool {
public:
    struct Ticket {
        std::atomic<size_t> remaining{0};
    };
    using TicketPtr = std::shared_ptr<Ticket>;

    explicit CopyPool(int nthreads)
    {
        for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this]() { worker(); });
    }

    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }

    int threads() const { return static_cast<int>(workers_.size()); }

    // Splits a (rows x row_bytes) strided copy into jobs of about `chunk` bytes and queues them.
    TicketPtr submit_2d(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t row_bytes, size_t rows,
                        bool streaming = true, size_t chunk = size_t(1) << 20)
    {
        auto t = std::make_shared<Ticket>();
        std::vector<Job> jobs;
        if (row_bytes == dst_pitch && row_bytes == src_pitch) {            // dense: one linear range
            const size_t total = row_bytes * rows;
            for (size_t o = 0; o < total; o += chunk) jobs.push_back({dst + o, src + o, std::min(chunk, total - o), 1, 0, 0, streaming, t});
        } else if (row_bytes >= chunk) {                                    // long rows: split each row
            for (size_t r = 0; r < rows; ++r)
                for (size_t o = 0; o < row_bytes; o += chunk)
                    jobs.push_back({dst + r * dst_pitch + o, src + r * src_pitch + o, std::min(chunk, row_bytes - o), 1, 0, 0, streaming, t});
        } else {                                                            // short rows: several rows per job
            const size_t per = std::max<size_t>(1, chunk / std::max<size_t>(row_bytes, 1));
            for (size_t r = 0; r < rows; r += per)
                jobs.push_back({dst + r * dst_pitch, src + r * src_pitch, row_bytes, std::min(per, rows - r), dst_pitch, src_pitch, streaming, t});
        }
        t->remaining.store(jobs.size(), std::memory_order_relaxed);
        if (jobs.empty()) return t;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (auto& j : jobs) q_.push_back(std::move(j));
        }
        cv_.notify_all();
        return t;
    }

    // Blocks until every job of the ticket has run; the caller executes queued jobs (its own or others') meanwhile.
    void wait(const TicketPtr& t)
    {
        if (!t) return;
        while (t->remaining.load(std::memory_order_acquire) != 0) {
            Job j;
            bool have = false;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (!q_.empty()) {
                    j = std::move(q_.front());
                    q_.pop_front();
                    have = true;
                }
            }
            if (have)
                run(j);
            else
                std::this_thread::yield();
        }
    }

private:
    struct Job {
        char* dst;
        const char* src;
        size_t bytes;        // per row
        size_t rows;
        size_t dst_pitch, src_pitch;
        bool streaming;      // non-temporal stores (default) or plain memcpy (the A/B arm: ctx option host_stage_nt=0)
        TicketPtr ticket;
    };

    static void run(Job& j)
    {
        for (size_t r = 0; r < j.rows; ++r) {
            if (j.streaming)
                stream_copy(j.dst + r * j.dst_pitch, j.src + r * j.src_pitch, j.bytes);
            else
                memcpy(j.dst + r * j.dst_pitch, j.src + r * j.src_pitch, j.bytes);
        }
        j.ticket->remaining.fetch_sub(1, std::memory_order_release);
    }

    void worker()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this]() { return stop_ || !q_.empty(); });
                if (q_.empty()) return;          // stop_ and drained
                j = std::move(q_.front());
                q_.pop_front();
            }
            run(j);
        }
    }

    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Job> q_;
    bool stop_ = false;
};

}  // namespace scir_b200
