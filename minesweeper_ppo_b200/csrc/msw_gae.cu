// msw_gae.cu -- RolloutBuffer.compute_gae (minesweeper/buffers.py:78-94) as one
// sm_100a kernel.
//
// The recurrence is serial in t and independent per env column, and the
// reference evaluates it in unfused IEEE fp32 in a fixed operation order, which
// this kernel reproduces bit for bit (__fmul_rn/__fadd_rn/__fsub_rn never
// contract into FMA):
//     nnt   = 1.0 - float(done[t])                                   (:89)
//     delta = (rewards[t] + (gamma * next_value) * nnt) - values[t]   (:90)
//     last  = delta + ((gamma*lam) * nnt) * last                      (:91)
//     returns = advantages + values                                   (:94)
// Storage is time-major [T][N] (buffers.py:81-83), so a row of 32 consecutive
// columns is one 128-byte line.  A CTA owns 16 or 32 columns: all 8 warps stream
// the [TC x COLS] tiles of rewards / values / dones into shared memory with
// coalesced loads (the only way to get enough bytes in flight when N is a few
// thousand columns).  Only `last` carries a dependency from t+1 to t -- next_value is
// values[t+1], plain data -- so all 256 threads then form delta[t] for their own
// elements in place, and the one warp that walks the chains does just
// FSEL / FMUL / FADD per step out of shared memory.  All warps write advantages /
// returns back coalesced.  HBM-bound in principle (17 B per transition) but
// latency-bound at the reference's sizes: forming delta inside the serial walk instead
// takes the same 14.3 us (A/B in one process, round 1), i.e. the walk is not the long pole.
#include "../../include/msw_b200.h"
#include "msw_error.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace msw {

constexpr int GAE_TC = 128;     // time rows per shared-memory chunk
constexpr int GAE_THREADS = 256;

// COLS env columns per CTA (16 or 32).  Thread t loads column t % COLS of rows t / COLS,
// t / COLS + 256 / COLS, ...; the first COLS lanes of warp 0 walk the chains.
template <int COLS>
__global__ void __launch_bounds__(GAE_THREADS)
gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
           const uint8_t *__restrict__ dones, const float *__restrict__ last_values,
           float *__restrict__ adv, float *__restrict__ ret, long long T, long long N,
           float gamma, float gamma_lam, int prescaled)
{
    __shared__ float s_r[GAE_TC][COLS];     // rewards in, then delta, then advantages
    __shared__ float s_v[GAE_TC][COLS];
    __shared__ uint8_t s_d[GAE_TC][COLS];
    __shared__ float s_gvn[COLS];           // gamma * value of the row after this chunk's last one
    constexpr int RPP = GAE_THREADS / COLS;       // rows per pass of the whole CTA
    constexpr int RPT = GAE_TC / RPP;             // rows per thread and chunk

    const int c = threadIdx.x % COLS;
    const int rr = threadIdx.x / COLS;
    const long long col = (long long)blockIdx.x * COLS + c;
    const bool in = col < N;
    const bool chain = threadIdx.x < COLS;        // warp 0, first COLS lanes

    float last = 0.0f;                                          // buffers.py:86
    if (chain) {                                                // buffers.py:88 (t == T-1)
        const float lv = in ? last_values[col] : 0.0f;
        // prescaled: last_values already is gamma*last_value (fp16 bootstrap)
        s_gvn[c] = prescaled ? lv : __fmul_rn(gamma, lv);
    }

    for (long long t_hi = T; t_hi > 0; t_hi -= GAE_TC) {
        const long long t_lo = t_hi > GAE_TC ? t_hi - GAE_TC : 0;
        const int rows = (int)(t_hi - t_lo);
        // ---- load: every thread issues all of its (up to 3*RPT) loads before the first use, so a
        // CTA has its whole tile in flight at once (the kernel is latency-, not bandwidth-bound)
        {
            float r[RPT], v[RPT];
            uint8_t d[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int i = rr + k * RPP;
                const bool ok = in && i < rows;
                const long long idx = (t_lo + i) * N + col;
                r[k] = ok ? __ldcs(rewards + idx) : 0.0f;
                v[k] = ok ? __ldcs(values + idx) : 0.0f;
                d[k] = ok ? __ldcs(dones + idx) : (uint8_t)0;
            }
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int i = rr + k * RPP;
                s_r[i][c] = r[k];
                s_v[i][c] = v[k];
                s_d[i][c] = d[k];
            }
        }
        __syncthreads();
        // ---- delta[t] = (r[t] + (gamma * v[t+1]) * nnt[t]) - v[t] for the thread's own elements
        // (buffers.py:88-90); each thread rewrites only the s_r entries it stored itself
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int i = rr + k * RPP;
            if (i < rows) {
                const float gv = i + 1 < rows ? __fmul_rn(gamma, s_v[i + 1][c]) : s_gvn[c];
                const float nn = s_d[i][c] ? 0.0f : 1.0f;
                s_r[i][c] = __fsub_rn(__fadd_rn(s_r[i][c], __fmul_rn(gv, nn)), s_v[i][c]);
            }
        }
        __syncthreads();
        // ---- chain: last = delta + ((gamma*lam) * nnt) * last (buffers.py:91), 8 rows per batch read
        // into registers ahead of the dependent arithmetic
        if (chain) {
            s_gvn[c] = __fmul_rn(gamma, s_v[0][c]);     // for the chunk before this one in time
            for (int hi = rows; hi > 0; hi -= 8) {
                float dl[8], gl[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = hi - 1 - k >= 0 ? hi - 1 - k : 0;
                    dl[k] = s_r[i][c];
                    gl[k] = s_d[i][c] ? 0.0f : gamma_lam;      // == gamma_lam * nnt exactly
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = hi - 1 - k;
                    if (i >= 0) {
                        last = __fadd_rn(dl[k], __fmul_rn(gl[k], last));
                        s_r[i][c] = last;
                    }
                }
            }
        }
        __syncthreads();
        if (in) {
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int i = rr + k * RPP;
                if (i < rows) {
                    const long long idx = (t_lo + i) * N + col;
                    const float a = s_r[i][c];
                    __stcs(adv + idx, a);
                    __stcs(ret + idx, __fadd_rn(a, s_v[i][c]));
                }
            }
        }
        __syncthreads();        // the next chunk overwrites the tiles and reads s_gvn
    }
}

// ---------------------------------------------------------------------------
// TMA variant (the default when N % 16 == 0 and the arrays are 16-byte aligned).  Measured on one B200
// (profiles/r01g_gae.md): the same 7.1 us as the plain-load kernel at C3 size (T=128, N=8,192: both are
// latency-bound, one tile per CTA), but 5.1 TB/s instead of 2.6 TB/s at N=524,288, where the kernel is
// HBM-bound.
//
// The tile loads and stores above cost 48 LDG + 48 STS + 32 STG per thread plus their address
// arithmetic, and the memory system sees them as ~100 k separate 128-byte requests.  Here one
// thread moves each [128 x 32] tile with a single cp.async.bulk.tensor.2d (rewards, values, dones
// in; advantages, returns out), completion on an mbarrier, so the CTA's whole 36 KB is in flight
// from the first cycle and no thread spends instructions on the copies.  Out-of-range rows and
// columns are zero-filled on load and clipped on store by the tensor map, so ragged T and N need
// no predicates.  Arithmetic and its order are exactly those of gae_kernel above.
// ---------------------------------------------------------------------------
constexpr int TG_ROWS = 128, TG_COLS = 32, TG_THREADS = 128;

struct TmaGaeSmem {
    alignas(128) float r[TG_ROWS][TG_COLS];      // rewards in, delta, advantages out
    alignas(128) float v[TG_ROWS][TG_COLS];      // values in, returns out
    alignas(128) float c[TG_ROWS][TG_COLS];      // (gamma*lam) * nnt
    alignas(128) uint8_t d[TG_ROWS][TG_COLS];
    float gvn[TG_COLS];                          // gamma * value of the row after this tile's last one
    alignas(8) unsigned long long bar;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}

__device__ __forceinline__ bool elect_one()      // one lane of a converged warp (elect.sync)
{
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0u;
}

__global__ void __launch_bounds__(TG_THREADS)
gae_tma_kernel(const __grid_constant__ CUtensorMap m_rewards, const __grid_constant__ CUtensorMap m_values,
               const __grid_constant__ CUtensorMap m_dones, const __grid_constant__ CUtensorMap m_adv,
               const __grid_constant__ CUtensorMap m_ret, const float *__restrict__ last_values, long long T,
               long long N, float gamma, float gamma_lam, int prescaled)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TmaGaeSmem &S = *reinterpret_cast<TmaGaeSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = blockIdx.x * TG_COLS;
    const bool chain = warp == 0;
    constexpr unsigned TILE_BYTES = TG_ROWS * TG_COLS * (4 + 4 + 1);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float last = 0.0f;                                          // buffers.py:86
    if (chain) {                                                // buffers.py:88 (t == T-1)
        const float lv = col0 + lane < N ? last_values[col0 + lane] : 0.0f;
        S.gvn[lane] = prescaled ? lv : __fmul_rn(gamma, lv);    // prescaled: already gamma*last_value (fp16 bootstrap)
    }
    __syncthreads();

    const long long tiles = (T + TG_ROWS - 1) / TG_ROWS;
    unsigned parity = 0;
    for (long long k = tiles - 1; k >= 0; --k, parity ^= 1u) {
        const int row0 = (int)(k * TG_ROWS);
        const int rows = (int)(T - row0 < TG_ROWS ? T - row0 : TG_ROWS);
        if (warp == 0 && elect_one()) {         // elect.sync: the TMA instructions issue without a divergence waterfall
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&S.bar)), "r"(TILE_BYTES) : "memory");
            tma_load_2d(&S.r[0][0], &m_rewards, col0, row0, &S.bar);
            tma_load_2d(&S.v[0][0], &m_values, col0, row0, &S.bar);
            tma_load_2d(&S.d[0][0], &m_dones, col0, row0, &S.bar);
        }
        {
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&S.bar)), "r"(parity) : "memory");
        }
        // ---- delta[t] = (r[t] + (gamma * v[t+1]) * nnt[t]) - v[t] (buffers.py:88-90) and c[t] =
        // (gamma*lam) * nnt[t]: warp w owns rows w, w+4, ..., lane = column
#pragma unroll 8
        for (int i = warp; i < rows; i += TG_THREADS / 32) {
            const float gv = i + 1 < rows ? __fmul_rn(gamma, S.v[i + 1][lane]) : S.gvn[lane];
            const bool dn = S.d[i][lane] != 0;
            S.r[i][lane] = __fsub_rn(__fadd_rn(S.r[i][lane], __fmul_rn(gv, dn ? 0.0f : 1.0f)), S.v[i][lane]);
            S.c[i][lane] = dn ? 0.0f : gamma_lam;               // == gamma_lam * nnt exactly
        }
        __syncthreads();
        // ---- chain: last = delta + c * last (buffers.py:91), 8 rows per batch read ahead of the dependent math
        if (chain) {
            S.gvn[lane] = __fmul_rn(gamma, S.v[0][lane]);       // for the tile before this one in time
            for (int hi = rows; hi > 0; hi -= 8) {
                float dl[8], gl[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = hi - 1 - q >= 0 ? hi - 1 - q : 0;
                    dl[q] = S.r[i][lane];
                    gl[q] = S.c[i][lane];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = hi - 1 - q;
                    if (i >= 0) {
                        last = __fadd_rn(dl[q], __fmul_rn(gl[q], last));
                        S.r[i][lane] = last;
                    }
                }
            }
        }
        __syncthreads();
        // ---- returns = advantages + values (buffers.py:94), in place over the values
#pragma unroll 8
        for (int i = warp; i < rows; i += TG_THREADS / 32) S.v[i][lane] = __fadd_rn(S.r[i][lane], S.v[i][lane]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the TMA store
        __syncthreads();
        if (warp == 0 && elect_one()) {
            tma_store_2d(&m_adv, col0, row0, &S.r[0][0]);
            tma_store_2d(&m_ret, col0, row0, &S.v[0][0]);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (k > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // tile buffers are reused
        }
        if (k > 0) __syncthreads();
    }
    if (warp == 0 && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static const EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [T][N] row-major array of `es`-byte elements, box = [TG_ROWS][TG_COLS]
static bool make_map(CUtensorMap *m, const void *base, CUtensorMapDataType dt, size_t es, int64_t T, int64_t N)
{
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
    const cuuint64_t strides[1] = {(cuuint64_t)N * es};
    const cuuint32_t box[2] = {TG_COLS, TG_ROWS}, estr[2] = {1, 1};
    return encode_tiled_fn()(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace msw

extern "C" int msw_gae(const float *rewards, const float *values, const uint8_t *dones,
                       const float *last_values, float *advantages, float *returns, int64_t T,
                       int64_t N, float gamma_f32, float gamma_lam_f32, int32_t last_values_prescaled,
                       void *stream)
{
    using namespace msw;
    if (!rewards || !values || !dones || !last_values || !advantages || !returns)
        return fail(MSW_ERR_NULL, "msw_gae: NULL pointer");
    if (T < 0 || N < 0) return fail(MSW_ERR_BAD_SHAPE, "msw_gae: T=%lld N=%lld", (long long)T, (long long)N);
    if (T == 0 || N == 0) return MSW_OK;
    // 32 columns per CTA (one full 128-byte line per row).  16 columns -- twice the CTAs, so every SM
    // has work at N = 8,192 -- was measured slower (18.4 vs 14.3 us, half-line requests) and removed.
    const int cols = 32;
    const long long blocks = (N + cols - 1) / cols;
    if (blocks > 0x7fffffffLL) return fail(MSW_ERR_BAD_SHAPE, "msw_gae: N too large");
    // TMA kernel where the tensor maps can be built (16-byte aligned bases and row pitches: N % 16 == 0, N bytes
    // for dones); the plain-load kernel serves ragged N.  One kernel per shape class, no run-time switch
    // (-DMSW_DEV_KNOBS: MSW_GAE_TMA=0 forces the plain kernel for tools/ measurements).
#ifdef MSW_DEV_KNOBS
    static const bool tma_on = [] {
        const char *e = getenv("MSW_GAE_TMA");
        return !(e && e[0] == '0');
    }();
#else
    constexpr bool tma_on = true;
#endif
    const uintptr_t bases = (uintptr_t)rewards | (uintptr_t)values | (uintptr_t)dones | (uintptr_t)advantages | (uintptr_t)returns;
    if (tma_on && N % 16 == 0 && (bases & 15u) == 0 && T <= 0x7fffffffLL && encode_tiled_fn()) {
        CUtensorMap mr, mv, md, ma, mt;
        if (!make_map(&mr, rewards, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N) ||
            !make_map(&mv, values, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N) ||
            !make_map(&md, dones, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, T, N) ||
            !make_map(&ma, advantages, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N) ||
            !make_map(&mt, returns, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, T, N))
            return fail(MSW_ERR_ARG, "msw_gae: cuTensorMapEncodeTiled failed (T=%lld N=%lld)", (long long)T, (long long)N);
        MSW_SET_MAX_SMEM(gae_tma_kernel, sizeof(TmaGaeSmem));
        gae_tma_kernel<<<(unsigned)blocks, TG_THREADS, sizeof(TmaGaeSmem), (cudaStream_t)stream>>>(
            mr, mv, md, ma, mt, last_values, T, N, gamma_f32, gamma_lam_f32, (int)last_values_prescaled);
        MSW_CUDA_TRY(cudaGetLastError());
        return MSW_OK;
    }
#define MSW_GAE_LAUNCH(C)                                                                           \
    gae_kernel<C><<<(unsigned)blocks, GAE_THREADS, 0, (cudaStream_t)stream>>>(                  \
        rewards, values, dones, last_values, advantages, returns, T, N, gamma_f32, gamma_lam_f32, \
        (int)last_values_prescaled)
    MSW_GAE_LAUNCH(32);
#undef MSW_GAE_LAUNCH
    MSW_CUDA_TRY(cudaGetLastError());
    return MSW_OK;
}
