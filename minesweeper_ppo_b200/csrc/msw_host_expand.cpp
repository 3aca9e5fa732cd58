// msw_host_expand.cpp -- HOST side of the reference calling convention (VecMinesweeper.step returning NumPy
// arrays, env.py:479-511): format conversion of the packed device state into the reference's fp32 observation
// planes (env.py:172-192) and bool action mask (env.py:194-196), multi-threaded on the host.  Pure host code
// (compiled by the host compiler; nvcc only drives the build).
//
// Why: the step itself runs on the GPU either way, but the reference-shaped result is 41*HW bytes per env
// (10.5 KB at 16x16) and crossing PCIe with it caps msw_step_host at ~5e6 env-steps/s (the speed of a 16-thread
// CPU port).  The state the observation is a pure function of -- the mine and revealed bitboards -- is
// 2*ceil(HW/32)*4 bytes per env (64 B at 16x16), so msw_step_host copies THAT device->host and expands it
// here, straight into the caller's (ordinary, unpinned) result arrays with non-temporal stores.  This is a
// format conversion of the GPU's result, not a CPU implementation of the env: no game logic runs here.
//
// Two modes.  FULL: every value of obs / mask is written (688 MB per step at 65,536 envs of 16x16: bound by the
// host's store bandwidth).  DELTA (msw_host_out.shadow / msw_expand_obs_host_delta): the caller owns the result
// arrays across calls and passes a `shadow` -- the bit planes those arrays currently hold, 10*HW bits per env --
// and only the 32- / 64-byte groups whose bits differ from the shadow are rewritten (one step changes ~16 of an
// env's 160 cache lines under random play, ~26 against the state two steps back), then the shadow is updated.
// The result is byte-identical to FULL by construction as long as nobody else wrote to the arrays.
//
// Counts are recomputed from the mine bitboard exactly as the device encoder does (bit-sliced adder over the
// eight neighbour rows; for 16x16 boards the whole board is one 256-bit AVX2 register, rows = 16-bit lanes);
// tests/test_host_expand.py checks both modes against the oracle's encoder on CPU and
// tests/test_gpu_env.py::test_numpy_api_is_reference_shaped end to end.
#include "../../include/msw_b200.h"
#include "msw_error.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <immintrin.h>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

namespace msw {

namespace {

// ---- a small persistent worker pool (the library owns no device memory; host threads are fine).
// One parallel region per msw_step_host call, ~1 ms of work: waking sleeping threads through a condition variable
// costs ~100 us of that, so an idle worker first SPINS on the region counter for SPIN_US (the next step's region
// normally arrives within that time when steps are issued back to back) and only then sleeps.
class HostPool {
public:
    static HostPool &get()
    {
        // intentionally leaked: workers may outlive static destruction.  A fork()ed child inherits the object but
        // none of its threads (a region would wait for workers that do not exist), so the child starts a fresh pool.
        static std::once_flag once;
        std::call_once(once, [] {
            instance().store(new HostPool(), std::memory_order_release);
            pthread_atfork(nullptr, nullptr, [] { instance().store(new HostPool(), std::memory_order_release); });
        });
        return *instance().load(std::memory_order_acquire);
    }
    // Runs fn(chunk) for chunk = 0..chunks-1 on up to `threads` threads (the caller is one of them).
    void run(int threads, long long chunks, const std::function<void(long long)> &fn)
    {
        if (threads < 1) threads = 1;
        if (threads > 4096) threads = 4096;
        if (chunks <= 1 || threads == 1) {
            for (long long c = 0; c < chunks; ++c) fn(c);
            return;
        }
        std::unique_lock<std::mutex> call_lock(call_mu_);     // one parallel region at a time
        grow(threads - 1);
        fn_ = &fn;
        chunks_ = chunks;
        next_.store(0, std::memory_order_relaxed);
        pending_.store(threads - 1, std::memory_order_relaxed);
        const unsigned long long gen = (state_.load(std::memory_order_relaxed) >> 16) + 1;
        {
            std::lock_guard<std::mutex> g(mu_);                // pairs with the sleepers' predicate check
            state_.store(gen << 16 | (unsigned long long)(threads - 1), std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_work_.notify_all();
        work();
        for (unsigned spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins) {
            if (spins < 20000) cpu_relax();
            else sched_yield();
        }
        fn_ = nullptr;
    }

private:
    static std::atomic<HostPool *> &instance()
    {
        static std::atomic<HostPool *> p{nullptr};
        return p;
    }
    static constexpr long long SPIN_US = 400;
    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        _mm_pause();
#endif
    }
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
        unsigned long long seen = 0;                          // region counter this worker has handled
        for (;;) {
            unsigned long long st = state_.load(std::memory_order_acquire);
            if ((st >> 16) == seen) {
                const auto t0 = std::chrono::steady_clock::now();
                for (unsigned spins = 1;; ++spins) {
                    cpu_relax();
                    st = state_.load(std::memory_order_acquire);
                    if ((st >> 16) != seen) break;
                    if ((spins & 127u) == 0 &&
                        std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > SPIN_US) {
                        std::unique_lock<std::mutex> g(mu_);
                        sleepers_.fetch_add(1, std::memory_order_acq_rel);
                        cv_work_.wait(g, [&] { return (state_.load(std::memory_order_acquire) >> 16) != seen; });
                        sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                        st = state_.load(std::memory_order_acquire);
                        break;
                    }
                }
            }
            seen = st >> 16;
            if (id >= (int)(st & 0xffffu)) continue;          // this region uses fewer threads
            work();
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_work_;
    std::vector<std::thread> workers_;
    const std::function<void(long long)> *fn_ = nullptr;
    std::atomic<long long> next_{0};
    std::atomic<unsigned long long> state_{0};                // region counter << 16 | workers taking part
    std::atomic<int> pending_{0}, sleepers_{0};
    long long chunks_ = 0;
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

constexpr int PLANE_WORDS_MAX = (MSW_OBS_CHANNELS * MSW_MAX_CELLS) / 64 + 2;

inline int shadow_words_for(int HW) { return (MSW_OBS_CHANNELS * HW + 63) / 64 + 1; }

// eight floats {0.0f, 1.0f}; STREAM = non-temporal (needs a 16-byte aligned destination)
template <bool STREAM>
inline void put8(float *dst, const float *src)
{
    if (STREAM) {
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst), _mm_load_si128(reinterpret_cast<const __m128i *>(src)));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 4), _mm_load_si128(reinterpret_cast<const __m128i *>(src + 4)));
    } else {
        memcpy(dst, src, 32);
    }
}

// The observation of one env is 10*HW values in {0, 1}: plane p, row r, column c = bit p*HW + r*W + c of a bit
// string that is assembled here from W-bit rows (`words` 64-bit words, the last one spare for put_bits).
inline void build_planes(const uint32_t *mines, const uint32_t *revealed, int first_click_done, int H, int W, int wpb,
                         uint64_t *planes, int words)
{
    const int HW = H * W;
    const uint32_t wmask = W == 32 ? 0xffffffffu : ((1u << W) - 1u);
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
}

// Bit string -> floats / mask bytes, front to back (one sequential stream per env, so non-temporal stores fill
// whole write-combining lines).  `old` (nullable) = the bit string the arrays hold now: groups of eight values
// whose bits are unchanged are skipped.
template <bool STREAM>
inline void emit_env(const uint64_t *planes, const uint64_t *old, int HW, float *obs, uint8_t *mask, const Luts &L)
{
    const uint8_t *nb = reinterpret_cast<const uint8_t *>(planes);
    const uint8_t *ob = reinterpret_cast<const uint8_t *>(old);
    if (obs) {
        const long long nbits = (long long)MSW_OBS_CHANNELS * HW, nfull = nbits >> 3;
        if (!old) {
            for (long long j = 0; j < nfull; ++j) put8<STREAM>(obs + 8 * j, L.f8[nb[j]]);
        } else {
            for (long long w = 0; w * 8 < nfull; ++w) {
                uint64_t x = planes[w] ^ old[w];
                while (x) {
                    const int k = __builtin_ctzll(x) >> 3;
                    x &= ~(0xffull << (8 * k));
                    const long long j = w * 8 + k;
                    if (j < nfull) put8<STREAM>(obs + 8 * j, L.f8[nb[j]]);
                }
            }
        }
        const int tail = (int)(nbits & 7);
        if (tail && (!old || ((nb[nfull] ^ ob[nfull]) & ((1u << tail) - 1u))))
            memcpy(obs + 8 * nfull, L.f8[nb[nfull]], (size_t)tail * 4);
    }
    if (mask) {
        // flat bit r*W+c of plane 0 (= the revealed bitboard) IS flat cell r*W+c of the mask: byte by byte
        const int mfull = HW >> 3, tail = HW & 7;
        for (int g = 0; g < mfull; ++g)
            if (!old || nb[g] != ob[g]) memcpy(mask + 8 * g, &L.b8[nb[g]], 8);
        if (tail && (!old || ((nb[mfull] ^ ob[mfull]) & ((1u << tail) - 1u))))
            memcpy(mask + 8 * mfull, &L.b8[nb[mfull]], (size_t)tail);
    }
}

template <bool STREAM>
void expand_range(const uint32_t *mines, const uint32_t *revealed, const int32_t *meta, long long lo, long long hi, int H,
                  int W, float *obs, uint8_t *mask, uint64_t *shadow, bool valid, const Luts &L)
{
    const int HW = H * W, wpb = (HW + 31) / 32, SW = shadow_words_for(HW);
    uint64_t planes[PLANE_WORDS_MAX];
    for (long long i = lo; i < hi; ++i) {
        build_planes(mines + i * wpb, revealed + i * wpb, meta ? meta[4 * i] : 1, H, W, wpb, planes, SW);
        uint64_t *sh = shadow ? shadow + i * SW : nullptr;
        emit_env<STREAM>(planes, (sh && valid) ? sh : nullptr, HW, obs ? obs + (size_t)i * MSW_OBS_CHANNELS * HW : nullptr,
                         mask ? mask + (size_t)i * HW : nullptr, L);
        if (sh) memcpy(sh, planes, (size_t)SW * 8);
    }
}

#if defined(__x86_64__)
#define MSW_HAVE_AVX2_PATH 1
// 16x16 boards (run-time dispatch on AVX2): the packed board is one 256-bit register whose 16-bit lane r is row r
// (bit c = column c), which is also plane p of the shadow / bit string (bits p*256 + 16 r + c).  Neighbour rows are
// whole-register byte shifts, neighbour columns are per-lane bit shifts (a 16-bit lane drops what leaves the row
// by itself); one changed lane = one 64-byte line of the fp32 plane.  (One 64-byte AVX-512 masked-move store per
// line instead of four 16-byte table stores was measured no faster: the scattered lines are bound by the core's
// write-combining buffers, ~9 GB/s per core, not by instruction issue.)
template <bool STREAM>
__attribute__((target("avx2"))) void expand16_range(const uint32_t *mines, const uint32_t *revealed, const int32_t *meta,
                                                     long long lo, long long hi, float *obs, uint8_t *mask, uint64_t *shadow,
                                                     bool valid, const Luts &L)
{
    constexpr int HW = 256, SW = (MSW_OBS_CHANNELS * HW + 63) / 64 + 1;
    const __m256i ones = _mm256_set1_epi32(-1);
    const bool delta = shadow && valid;
    for (long long i = lo; i < hi; ++i) {
        const __m256i R = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(revealed + i * 8));
        __m256i pl[MSW_OBS_CHANNELS];
        pl[0] = R;
        if ((!meta || meta[4 * i]) && !_mm256_testz_si256(R, R)) {                                       // env.py:181
            const __m256i M = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(mines + i * 8));
            const __m256i up = _mm256_alignr_epi8(M, _mm256_permute2x128_si256(M, M, 0x08), 14);   // lane r = row r-1
            const __m256i dn = _mm256_alignr_epi8(_mm256_permute2x128_si256(M, M, 0x81), M, 2);    // lane r = row r+1
            const __m256i a = _mm256_slli_epi16(up, 1), b = up, c = _mm256_srli_epi16(up, 1);
            const __m256i d = _mm256_slli_epi16(M, 1), e = _mm256_srli_epi16(M, 1);
            const __m256i f = _mm256_slli_epi16(dn, 1), g = dn, h = _mm256_srli_epi16(dn, 1);
#define X_(p, q) _mm256_xor_si256(p, q)
#define A_(p, q) _mm256_and_si256(p, q)
#define O_(p, q) _mm256_or_si256(p, q)
            const __m256i ab = X_(a, b), s0 = X_(ab, c), c0 = O_(A_(a, b), A_(c, ab));
            const __m256i de = X_(d, e), s1 = X_(de, f), c1 = O_(A_(d, e), A_(f, de));
            const __m256i s2 = X_(g, h), c2 = A_(g, h);
            const __m256i s01 = X_(s0, s1), n0 = X_(s01, s2), c3 = O_(A_(s0, s1), A_(s2, s01));
            const __m256i c01 = X_(c0, c1), t = X_(c01, c2), c4 = O_(A_(c0, c1), A_(c2, c01));
            const __m256i n1 = X_(t, c3), c5 = A_(t, c3);
            const __m256i n2 = X_(c4, c5), n3 = A_(c4, c5);
            const __m256i bit[4][2] = {{X_(n0, ones), n0}, {X_(n1, ones), n1}, {X_(n2, ones), n2}, {X_(n3, ones), n3}};
            for (int k = 0; k < 9; ++k)                                                       // one-hot count at revealed cells
                pl[1 + k] = A_(A_(A_(bit[0][k & 1], bit[1][k >> 1 & 1]), A_(bit[2][k >> 2 & 1], bit[3][k >> 3 & 1])), R);
#undef X_
#undef A_
#undef O_
        } else {
            for (int k = 1; k < MSW_OBS_CHANNELS; ++k) pl[k] = _mm256_setzero_si256();
        }
        uint64_t *sh = shadow ? shadow + i * SW : nullptr;
        if (mask) {
            uint32_t cm = 0xffffffffu;                                                        // groups of 8 cells to rewrite
            if (delta)
                cm = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(R, _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sh))));
            const uint8_t *rb = reinterpret_cast<const uint8_t *>(revealed + i * 8);
            uint8_t *mk = mask + (size_t)i * HW;
            while (cm) {
                const int g = __builtin_ctz(cm);
                cm &= cm - 1;
                memcpy(mk + 8 * g, &L.b8[rb[g]], 8);
            }
        }
        float *ob = obs ? obs + (size_t)i * MSW_OBS_CHANNELS * HW : nullptr;
        for (int p = 0; p < MSW_OBS_CHANNELS; ++p) {
            uint32_t ch = 0xffffffffu;                                                        // two bits per changed row
            if (delta)
                ch = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi16(pl[p], _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sh + 4 * p))));
            if (!ch) continue;
            if (ob) {
                alignas(32) uint16_t rows[16];
                _mm256_store_si256(reinterpret_cast<__m256i *>(rows), pl[p]);
                float *dst = ob + p * HW;
                while (ch) {
                    const int r = __builtin_ctz(ch) >> 1;
                    ch &= ~(3u << (2 * r));
                    put8<STREAM>(dst + 16 * r, L.f8[rows[r] & 255]);
                    put8<STREAM>(dst + 16 * r + 8, L.f8[rows[r] >> 8]);
                }
            }
            if (sh) _mm256_storeu_si256(reinterpret_cast<__m256i *>(sh + 4 * p), pl[p]);
        }
        if (sh && !delta) sh[SW - 1] = 0;
    }
}
#endif

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

// mines / revealed: [n][wpb] words, meta: [n][4] (word 0 = first_click_done; nullable: a cell can only be revealed
// after the first click, so first_click_done is implied wherever it matters); obs [n][10][H][W] / mask [n][HW] nullable.
// shadow (nullable): [n][msw_shadow_words(H, W)] -- the bit planes obs / mask hold; shadow_valid = 0: the arrays
// hold anything (everything is written and the shadow initialised), != 0: only what differs is rewritten.
void expand_obs_host(int H, int W, const uint32_t *mines, const uint32_t *revealed, const int32_t *meta, long long n,
                     float *obs, uint8_t *mask, uint64_t *shadow, int shadow_valid, int threads)
{
    const int HW = H * W;
    const Luts &L = luts();
    // streaming stores need every group of eight floats 16-byte aligned: base aligned and 10*HW floats per env a
    // multiple of 16 bytes (256-bit stores were measured no faster: one core sustains ~7.5 GB/s of non-temporal
    // stores either way)
    const bool stream = obs && (((uintptr_t)obs & 15u) == 0) && ((MSW_OBS_CHANNELS * HW) % 4 == 0);
    const bool valid = shadow && shadow_valid;
#ifdef MSW_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const bool fast16 = avx2 && H == 16 && W == 16;
#else
    const bool fast16 = false;
#endif
    const long long chunk = 128;
    const long long chunks = (n + chunk - 1) / chunk;
    const std::function<void(long long)> fn = [&](long long c) {
        const long long lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
#ifdef MSW_HAVE_AVX2_PATH
        if (fast16) {
            if (stream) expand16_range<true>(mines, revealed, meta, lo, hi, obs, mask, shadow, valid, L);
            else        expand16_range<false>(mines, revealed, meta, lo, hi, obs, mask, shadow, valid, L);
        } else
#endif
        if (stream) expand_range<true>(mines, revealed, meta, lo, hi, H, W, obs, mask, shadow, valid, L);
        else        expand_range<false>(mines, revealed, meta, lo, hi, H, W, obs, mask, shadow, valid, L);
        if (stream) _mm_sfence();
    };
    HostPool::get().run(host_thread_count(threads), chunks, fn);
}

}  // namespace msw

static int check_expand_args(const msw_env_desc *desc, const void *a, const void *b, const void *c, int64_t n, const char *who)
{
    using namespace msw;
    (void)c;                                           // h_meta may be NULL
    if (!desc || !a || !b) return fail(MSW_ERR_NULL, "%s: NULL pointer", who);
    if (msw_words_per_board(desc->H, desc->W) == 0)
        return fail(MSW_ERR_BAD_SHAPE, "%s: unsupported board %dx%d", who, desc->H, desc->W);
    if (n < 0) return fail(MSW_ERR_BAD_SHAPE, "%s: n=%lld", who, (long long)n);
    return MSW_OK;
}

extern "C" int msw_shadow_words(int32_t H, int32_t W)
{
    if (msw_words_per_board(H, W) == 0) return 0;
    return msw::shadow_words_for(H * W);
}

extern "C" int msw_expand_obs_host(const msw_env_desc *desc, const int32_t *h_mines, const int32_t *h_revealed,
                                   const int32_t *h_meta, int64_t n, float *h_obs, uint8_t *h_mask, int32_t threads)
{
    using namespace msw;
    const int rc = check_expand_args(desc, h_mines, h_revealed, h_meta, n, "msw_expand_obs_host");
    if (rc) return rc;
    if (n == 0 || (!h_obs && !h_mask)) return MSW_OK;
    expand_obs_host(desc->H, desc->W, reinterpret_cast<const uint32_t *>(h_mines), reinterpret_cast<const uint32_t *>(h_revealed),
                    h_meta, (long long)n, h_obs, h_mask, nullptr, 0, threads);
    return MSW_OK;
}

extern "C" int msw_expand_obs_host_delta(const msw_env_desc *desc, const int32_t *h_mines, const int32_t *h_revealed,
                                         const int32_t *h_meta, int64_t n, float *h_obs, uint8_t *h_mask,
                                         uint64_t *shadow, int32_t shadow_valid, int32_t threads)
{
    using namespace msw;
    const int rc = check_expand_args(desc, h_mines, h_revealed, h_meta, n, "msw_expand_obs_host_delta");
    if (rc) return rc;
    if (!h_obs || !h_mask || !shadow)
        return fail(MSW_ERR_NULL, "msw_expand_obs_host_delta: obs, mask and shadow are all required (the shadow describes both arrays)");
    if (n == 0) return MSW_OK;
    expand_obs_host(desc->H, desc->W, reinterpret_cast<const uint32_t *>(h_mines), reinterpret_cast<const uint32_t *>(h_revealed),
                    h_meta, (long long)n, h_obs, h_mask, shadow, shadow_valid, threads);
    return MSW_OK;
}
