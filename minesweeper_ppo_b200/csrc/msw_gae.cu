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
    // has work at N = 8,192 -- was measured slower (18.4 vs 14.3 us, half-line requests); it stays
    // selectable through MSW_GAE_COLS=16 for experiments.
    static const int forced = [] {
        const char *e = getenv("MSW_GAE_COLS");
        return e ? atoi(e) : 0;
    }();
    const int cols = forced == 16 ? 16 : 32;
    const long long blocks = (N + cols - 1) / cols;
    if (blocks > 0x7fffffffLL) return fail(MSW_ERR_BAD_SHAPE, "msw_gae: N too large");
#define MSW_GAE_LAUNCH(C)                                                                           \
    gae_kernel<C><<<(unsigned)blocks, GAE_THREADS, 0, (cudaStream_t)stream>>>(                  \
        rewards, values, dones, last_values, advantages, returns, T, N, gamma_f32, gamma_lam_f32, \
        (int)last_values_prescaled)
    if (cols == 16) MSW_GAE_LAUNCH(16);
    else            MSW_GAE_LAUNCH(32);
#undef MSW_GAE_LAUNCH
    MSW_CUDA_TRY(cudaGetLastError());
    return MSW_OK;
}
