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
// columns is one 128-byte line.  A CTA owns 32 columns: all 8 warps stream the
// [TC x 32] tiles of rewards / values / dones into shared memory with coalesced
// loads (the only way to get enough bytes in flight when N is a few thousand
// columns), warp 0 walks the chain out of shared memory, and all warps write
// advantages / returns back coalesced.  HBM-bound in principle (17 B per
// transition) but latency-bound at the reference's sizes.
#include "../../include/msw_b200.h"
#include "msw_error.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace msw {

constexpr int GAE_COLS = 32;    // columns per CTA (one 128 B line per row)
constexpr int GAE_TC = 128;     // time rows per shared-memory chunk
constexpr int GAE_WARPS = 8;

__global__ void __launch_bounds__(GAE_WARPS * 32)
gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
           const uint8_t *__restrict__ dones, const float *__restrict__ last_values,
           float *__restrict__ adv, float *__restrict__ ret, long long T, long long N,
           float gamma, float gamma_lam, int prescaled)
{
    __shared__ float s_r[GAE_TC][GAE_COLS];     // rewards in, advantages out
    __shared__ float s_v[GAE_TC][GAE_COLS];
    __shared__ uint8_t s_d[GAE_TC][GAE_COLS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * GAE_COLS + lane;
    const bool in = col < N;
    constexpr int RPW = GAE_TC / GAE_WARPS;     // rows per warp and chunk (16)

    float last = 0.0f;                                          // buffers.py:86
    float next_value = in ? last_values[col] : 0.0f;            // buffers.py:88 (t == T-1)
    bool scaled = prescaled != 0;       // last_values already is gamma*last_value (fp16 bootstrap)

    for (long long t_hi = T; t_hi > 0; t_hi -= GAE_TC) {
        const long long t_lo = t_hi > GAE_TC ? t_hi - GAE_TC : 0;
        const int rows = (int)(t_hi - t_lo);
        __syncthreads();
        // ---- load: every thread issues all of its (up to 3*RPW) loads before the first use, so a
        // CTA has its whole 36 KB tile in flight at once (the kernel is latency-, not bandwidth-bound)
        {
            float r[RPW], v[RPW];
            uint8_t d[RPW];
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const int i = warp + k * GAE_WARPS;
                const bool ok = in && i < rows;
                const long long idx = (t_lo + i) * N + col;
                r[k] = ok ? __ldcs(rewards + idx) : 0.0f;
                v[k] = ok ? __ldcs(values + idx) : 0.0f;
                d[k] = ok ? __ldcs(dones + idx) : (uint8_t)0;
            }
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const int i = warp + k * GAE_WARPS;
                s_r[i][lane] = r[k];
                s_v[i][lane] = v[k];
                s_d[i][lane] = d[k];
            }
        }
        __syncthreads();
        // ---- chain: warp 0, 8 rows per batch read into registers ahead of the dependent arithmetic
        if (warp == 0) {
            for (int hi = rows; hi > 0; hi -= 8) {
                float r[8], v[8], nn[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = hi - 1 - k;
                    const bool ok = i >= 0;
                    r[k] = ok ? s_r[ok ? i : 0][lane] : 0.0f;
                    v[k] = ok ? s_v[ok ? i : 0][lane] : 0.0f;
                    nn[k] = (ok && s_d[ok ? i : 0][lane]) ? 0.0f : 1.0f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = hi - 1 - k;
                    if (i >= 0) {
                        const float gv = scaled ? next_value : __fmul_rn(gamma, next_value);
                        scaled = false;
                        const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(gv, nn[k])), v[k]);
                        last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lam, nn[k]), last));
                        s_r[i][lane] = last;
                        next_value = v[k];
                    }
                }
            }
        }
        __syncthreads();
        if (in) {
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const int i = warp + k * GAE_WARPS;
                if (i < rows) {
                    const long long idx = (t_lo + i) * N + col;
                    const float a = s_r[i][lane];
                    __stcs(adv + idx, a);
                    __stcs(ret + idx, __fadd_rn(a, s_v[i][lane]));
                }
            }
        }
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
    const long long blocks = (N + GAE_COLS - 1) / GAE_COLS;
    if (blocks > 0x7fffffffLL) return fail(MSW_ERR_BAD_SHAPE, "msw_gae: N too large");
    gae_kernel<<<(unsigned)blocks, GAE_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rewards, values, dones, last_values, advantages, returns, T, N, gamma_f32, gamma_lam_f32,
        (int)last_values_prescaled);
    MSW_CUDA_TRY(cudaGetLastError());
    return MSW_OK;
}
