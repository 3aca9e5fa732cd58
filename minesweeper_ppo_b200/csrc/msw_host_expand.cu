// msw_host_expand.cu -- HOST side of the reference calling convention (VecMinesweeper.step returning NumPy
// arrays, env.py:479-511): format conversion of the packed device state into the reference's fp32 observation
// planes (env.py:172-192) and bool action mask (env.py:194-196), multi-threaded on the host.
//
// Why: the step itself runs on the GPU either way, but the reference-shaped result is 41*HW bytes per env
// (10.5 KB at 16x16) and crossing PCIe with it caps msw_step_host at ~5e6 env-steps/s (the speed of a 16-thread
// CPU port).  The state the observation is a pure function of -- mines, revealed, first_click_done -- is
// 2*ceil(HW/32)*4 + 16 bytes per env (80 B at 16x16), so msw_step_host copies THAT device->host and expands it
// here, straight into the caller's (ordinary, unpinned) result arrays with non-temporal stores.  This is a
// format conversion of the GPU's result, not a CPU implementation of the env: no game logic runs here.
//
// Counts are recomputed from the mine bitboard exactly as the device encoder does (bit-sliced adder over the
// eight neighbour rows); tests/test_host_expand.py checks the expansion against the oracle's encoder on CPU
// and tests/test_gpu_env.py::test_numpy_api_is_reference_shaped end to end.
#include "../../include/msw_b200.h"
#include "msw_error.h"

#include <atomic>
#include <condition_variable>
#include <emmintrin.h>
#include <functional>
#include <mutex>
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

namespace msw {

namespace {

// ---- a small persistent worker pool (the library owns no device memory; host threads are fine)
class HostPool {
public:
    static HostPool &get()
    {
        static HostPool *p = new HostPool();      // intentionally leaked: workers may outlive static destruction
        return *p;
    }
    // Runs fn(chunk) for chunk = 0..chunks-1 on up to `threads` threads (the caller is one of them).
    void run(int threads, long long chunks, const std::function<void(long long)> &fn)
    {
        if (threads < 1) threads = 1;
        if (chunks <= 1 || threads == 1) {
            for (long long c = 0; c < chunks; ++c) fn(c);
            return;
        }
        std::unique_lock<std::mutex> call_lock(call_mu_);     // one parallel region at a time
        grow(threads - 1);
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn;
            chunks_ = chunks;
            next_.store(0, std::memory_order_relaxed);
            active_ = threads - 1;
            pending_ = active_;
            ++generation_;
        }
        cv_work_.notify_all();
        work();
        std::unique_lock<std::mutex> g(mu_);
        cv_done_.wait(g, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work()
    {
        for (;;) {
            const long long c = next_.fetch_add(1, std::memory_order_relaxed);
            if (c >= chunks_) break;
            (*fn_)(c);
        }
    }
    void grow(int want)
    {
        while ((int)workers_.size() < want) {
            const int id = (int)workers_.size();
            workers_.emplace_back([this, id] { loop(id); });
            workers_.back().detach();
        }
    }
    void loop(int id)
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_work_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                if (id >= active_) continue;                  // this region uses fewer threads
            }
            work();
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    const std::function<void(long long)> *fn_ = nullptr;
    std::atomic<long long> next_{0};
    long long chunks_ = 0;
    int active_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
};

struct Luts {
    alignas(64) float f8[256][8];       // byte -> eight floats {0.0f, 1.0f}
    alignas(64) uint64_t b8[256];       // byte -> eight bool bytes, INVERTED (mask = ~revealed)
    Luts()
    {
        for (int v = 0; v < 256; ++v) {
            uint64_t m = 0;
            for (int k = 0; k < 8; ++k) {
                f8[v][k] = (v >> k & 1) ? 1.0f : 0.0f;
                if (!(v >> k & 1)) m |= 1ull << (8 * k);
            }
            b8[v] = m;
        }
    }
};
const Luts &luts()
{
    static const Luts l;
    return l;
}

inline uint32_t row_bits(const uint32_t *b, int wpb, int r, int W, uint32_t wmask)
{
    const int bit = r * W, w = bit >> 5, s = bit & 31;
    const uint64_t lo = b[w], hi = (w + 1 < wpb) ? b[w + 1] : 0;
    return (uint32_t)(((hi << 32) | lo) >> s) & wmask;
}

// OR the low `w` bits of `v` into the bit string `dst` at bit offset `pos`
inline void put_bits(uint64_t *dst, long long pos, uint32_t v, int w)
{
    const long long word = pos >> 6;
    const int s = (int)(pos & 63);
    dst[word] |= (uint64_t)v << s;
    if (s + w > 64) dst[word + 1] |= (uint64_t)v >> (64 - s);
}

// `nbits` bits -> `nbits` floats {0.0f, 1.0f}, written front to back (one sequential stream per env, so
// non-temporal stores fill whole write-combining lines)
template <bool STREAM>
inline void put_floats(float *dst, const uint64_t *bits, long long nbits, const Luts &L)
{
    const uint8_t *by = reinterpret_cast<const uint8_t *>(bits);
    long long i = 0;
    for (; i + 8 <= nbits; i += 8) {
        const float *src = L.f8[by[i >> 3]];
        if (STREAM) {
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_load_si128(reinterpret_cast<const __m128i *>(src)));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm_load_si128(reinterpret_cast<const __m128i *>(src + 4)));
        } else {
            memcpy(dst + i, src, 32);
        }
    }
    if (i < nbits) memcpy(dst + i, L.f8[by[i >> 3]], (size_t)(nbits - i) * 4);
}

// The observation of one env is 10*HW values in {0, 1}: plane p, row r, column c = bit p*HW + r*W + c of a bit
// string that is assembled here from W-bit rows and then expanded to floats in one sequential pass.
template <bool STREAM>     // non-temporal stores (needs a 16-byte aligned destination)
void expand_env(const uint32_t *mines, const uint32_t *revealed, int first_click_done, int H, int W, int wpb,
                float *obs, uint8_t *mask, const Luts &L)
{
    const int HW = H * W;
    const uint32_t wmask = W == 32 ? 0xffffffffu : ((1u << W) - 1u);
    if (obs) {
        uint64_t planes[(MSW_OBS_CHANNELS * MSW_MAX_CELLS) / 64 + 2];
        const int words = (MSW_OBS_CHANNELS * HW + 63) / 64 + 1;
        for (int w = 0; w < words; ++w) planes[w] = 0;
        uint32_t up = 0, mid = row_bits(mines, wpb, 0, W, wmask);
        for (int r = 0; r < H; ++r) {
            const uint32_t dn = r + 1 < H ? row_bits(mines, wpb, r + 1, W, wmask) : 0u;
            const uint32_t v = row_bits(revealed, wpb, r, W, wmask);
            put_bits(planes, (long long)r * W, v, W);                                             // channel 0: revealed
            const uint32_t vis = first_click_done ? v : 0u;                                       // env.py:181
            if (vis) {
                // 8-neighbour count of every cell of the row as four bit planes (carry-save adder tree)
                const uint32_t a = (up << 1) & wmask, b = up, c = up >> 1, d = (mid << 1) & wmask, e = mid >> 1,
                               f = (dn << 1) & wmask, g = dn, h = dn >> 1;
                const uint32_t s0 = a ^ b ^ c, c0 = (a & b) | (c & (a ^ b));
                const uint32_t s1 = d ^ e ^ f, c1 = (d & e) | (f & (d ^ e));
                const uint32_t s2 = g ^ h, c2 = g & h;
                const uint32_t n0 = s0 ^ s1 ^ s2, c3 = (s0 & s1) | (s2 & (s0 ^ s1));
                const uint32_t t = c0 ^ c1 ^ c2, c4 = (c0 & c1) | (c2 & (c0 ^ c1));
                const uint32_t n1 = t ^ c3, c5 = t & c3;
                const uint32_t n2 = c4 ^ c5, n3 = c4 & c5;
                for (int k = 0; k < 9; ++k) {                                                     // channels 1..9: one-hot count
                    const uint32_t eq = ((k & 1) ? n0 : ~n0) & ((k & 2) ? n1 : ~n1) & ((k & 4) ? n2 : ~n2) & ((k & 8) ? n3 : ~n3);
                    const uint32_t bits = vis & eq;
                    if (bits) put_bits(planes, (long long)(1 + k) * HW + (long long)r * W, bits, W);
                }
            }
            up = mid;
            mid = dn;
        }
        put_floats<STREAM>(obs, planes, (long long)MSW_OBS_CHANNELS * HW, L);
    }
    if (mask) {
        // flat bit r*W+c of the bitboard IS flat cell r*W+c of the mask: byte by byte
        const uint8_t *rb = reinterpret_cast<const uint8_t *>(revealed);
        int cell = 0;
        for (; cell + 8 <= HW; cell += 8) memcpy(mask + cell, &L.b8[rb[cell >> 3]], 8);
        if (cell < HW) memcpy(mask + cell, &L.b8[rb[cell >> 3]], (size_t)(HW - cell));
    }
}

}  // namespace

int host_thread_count(int requested)
{
    if (requested > 0) return requested;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

// mines / revealed: [n][wpb] words, meta: [n][4] (word 0 = first_click_done); obs [n][10][H][W] / mask [n][HW] nullable
void expand_obs_host(int H, int W, const uint32_t *mines, const uint32_t *revealed, const int32_t *meta, long long n,
                     float *obs, uint8_t *mask, int threads)
{
    const int HW = H * W, wpb = (HW + 31) / 32;
    const Luts &L = luts();
    // streaming stores need every env's block 16-byte aligned: base aligned and 10*HW floats a multiple of 16 bytes
    // (256-bit AVX2 stores were measured no faster: one core sustains ~7.5 GB/s of non-temporal stores either way)
    const bool stream = obs && (((uintptr_t)obs & 15u) == 0) && ((MSW_OBS_CHANNELS * HW) % 4 == 0);
    const long long chunk = 128;
    const long long chunks = (n + chunk - 1) / chunk;
    const std::function<void(long long)> fn = [&](long long c) {
        const long long lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        for (long long i = lo; i < hi; ++i) {
            float *o = obs ? obs + (size_t)i * MSW_OBS_CHANNELS * HW : nullptr;
            uint8_t *m = mask ? mask + (size_t)i * HW : nullptr;
            if (stream) expand_env<true>(mines + i * wpb, revealed + i * wpb, meta[4 * i], H, W, wpb, o, m, L);
            else        expand_env<false>(mines + i * wpb, revealed + i * wpb, meta[4 * i], H, W, wpb, o, m, L);
        }
        if (stream) _mm_sfence();
    };
    HostPool::get().run(host_thread_count(threads), chunks, fn);
}

}  // namespace msw

extern "C" int msw_expand_obs_host(const msw_env_desc *desc, const int32_t *h_mines, const int32_t *h_revealed,
                                   const int32_t *h_meta, int64_t n, float *h_obs, uint8_t *h_mask, int32_t threads)
{
    using namespace msw;
    if (!desc || !h_mines || !h_revealed || !h_meta) return fail(MSW_ERR_NULL, "msw_expand_obs_host: NULL pointer");
    if (msw_words_per_board(desc->H, desc->W) == 0)
        return fail(MSW_ERR_BAD_SHAPE, "msw_expand_obs_host: unsupported board %dx%d", desc->H, desc->W);
    if (n < 0) return fail(MSW_ERR_BAD_SHAPE, "msw_expand_obs_host: n=%lld", (long long)n);
    if (n == 0 || (!h_obs && !h_mask)) return MSW_OK;
    expand_obs_host(desc->H, desc->W, reinterpret_cast<const uint32_t *>(h_mines), reinterpret_cast<const uint32_t *>(h_revealed),
                    h_meta, (long long)n, h_obs, h_mask, threads);
    return MSW_OK;
}
