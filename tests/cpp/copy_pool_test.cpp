// CPU test of the host-side staging machinery of the *_host entry points (scir_b200/csrc/copy_pool.hpp): the
// non-temporal stream_copy for every alignment / length class and CopyPool::submit_2d for dense, long-row and
// short-row strided shapes with several worker threads and overlapping submissions.  No GPU involved.
#include "../../scir_b200/csrc/copy_pool.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

static int failures = 0;
#define EXPECT(c)                                                          \
    do {                                                                   \
        if (!(c)) {                                                        \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);       \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

using scir_b200::CopyPool;

int main()
{
    // stream_copy: all destination / source misalignments, lengths around the 4096-byte and 64-byte thresholds
    {
        std::vector<unsigned char> src(1 << 16), dst(1 << 16);
        for (size_t i = 0; i < src.size(); ++i) src[i] = static_cast<unsigned char>((i * 131u + 7u) & 0xff);
        const size_t lens[] = {0, 1, 63, 64, 65, 4095, 4096, 4097, 4096 + 63, 10000, 40000};
        for (size_t so = 0; so < 70; so += 7)
            for (size_t dofs = 0; dofs < 70; dofs += 5)
                for (size_t len : lens) {
                    std::fill(dst.begin(), dst.end(), 0xEE);
                    scir_b200::stream_copy(reinterpret_cast<char*>(dst.data()) + dofs, reinterpret_cast<const char*>(src.data()) + so, len);
                    bool ok = true;
                    for (size_t i = 0; i < len && ok; ++i) ok = dst[dofs + i] == src[so + i];
                    ok = ok && (dofs == 0 || dst[dofs - 1] == 0xEE) && dst[dofs + len] == 0xEE;     // no byte outside the range touched
                    EXPECT(ok);
                }
    }
    // submit_2d: (rows, row_bytes, src_pitch, dst_pitch, chunk) -- dense, long rows split, short rows batched
    {
        CopyPool pool(4);
        EXPECT(pool.threads() == 4);
        struct Case { size_t rows, row_bytes, sp, dp, chunk; };
        const Case cases[] = {{1, 1 << 20, 1 << 20, 1 << 20, 1 << 16}, {37, 4096, 4096, 4096, 10000}, {5, 300000, 300016, 300032, 1 << 16},
                              {1000, 120, 128, 136, 4096}, {3, 1, 5, 7, 64}, {0, 100, 100, 100, 64}, {9, 0, 16, 16, 64}};
        for (int streaming = 0; streaming < 2; ++streaming)
            for (const Case& c : cases) {
                std::vector<unsigned char> src(c.rows * c.sp + 64), dst(c.rows * c.dp + 64, 0xEE);
                for (size_t i = 0; i < src.size(); ++i) src[i] = static_cast<unsigned char>((i * 29u + 3u) & 0xff);
                auto t = pool.submit_2d(reinterpret_cast<char*>(dst.data()), c.dp, reinterpret_cast<const char*>(src.data()), c.sp, c.row_bytes,
                                        c.rows, streaming != 0, c.chunk);
                pool.wait(t);
                bool ok = true;
                for (size_t r = 0; r < c.rows && ok; ++r) {
                    for (size_t i = 0; i < c.row_bytes && ok; ++i) ok = dst[r * c.dp + i] == src[r * c.sp + i];
                    for (size_t i = c.row_bytes; i < c.dp && ok; ++i) ok = dst[r * c.dp + i] == 0xEE;   // the pitch gap stays untouched
                }
                EXPECT(ok);
            }
        // several submissions in flight, waited out of order (what host_pipeline does with its in / out tickets)
        std::vector<std::vector<unsigned char>> srcs(8, std::vector<unsigned char>(1 << 20)), dsts(8, std::vector<unsigned char>(1 << 20, 0));
        std::vector<CopyPool::TicketPtr> tickets;
        for (int i = 0; i < 8; ++i) {
            for (size_t j = 0; j < srcs[i].size(); ++j) srcs[i][j] = static_cast<unsigned char>((j + 17u * i) & 0xff);
            tickets.push_back(pool.submit_2d(reinterpret_cast<char*>(dsts[i].data()), 1 << 20, reinterpret_cast<const char*>(srcs[i].data()), 1 << 20,
                                             1 << 20, 1));
        }
        for (int i = 7; i >= 0; --i) {
            pool.wait(tickets[i]);
            EXPECT(dsts[i] == srcs[i]);
        }
        pool.wait(nullptr);                                                // a null ticket is a no-op
    }
    std::printf("copy_pool: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
