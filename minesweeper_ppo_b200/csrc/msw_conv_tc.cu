// msw_conv_tc.cu -- 3x3 "same" convolution of the residual trunk (96 -> 96 channels on 16x16 boards,
// fp16 NHWC in, fp32 accumulate, fp16 NHWC out, no bias: the bias is folded into msw_gn_act) on the
// 5th-generation tensor cores.
//
// out[y][x][co] = sum_{dy,dx,ci} in[y+dy][x+dx][ci] * w[co][ci][dy][dx] is computed as nine shifted GEMMs
// per 128-pixel tile (8 image rows of one board), with the two kinds of shift handled where they are free:
//   * vertical: the tile's input rows y0-1 .. y0+8 sit in shared memory as ONE linear [160 pixel][96 ch]
//     block (a single 4-D TMA box; rows above / below the board are zero-filled by the tensor map), so the A
//     operand of tap dy is the same block at a byte offset of (dy+1) * 16 pixels -- a multiple of the swizzle
//     period, i.e. just another shared-memory descriptor;
//   * horizontal: tap dx accumulates into its own TMEM accumulator D_dx WITHOUT shifting the pixels, and the
//     epilogue forms out[y][x] = D_-1[y][x-1] + D_0[y][x] + D_+1[y][x+1].  A thread owns one pixel (TMEM lane),
//     a warp owns two image rows, so x-1 / x+1 are the neighbouring lanes (one shuffle each) and the board
//     edge coincides with the lanes that have no neighbour.
// All nine [96 x 96] weight taps stay resident in shared memory (162 KB, loaded once per persistent CTA);
// the activation streams through a 2-stage ring.  K = 96 = one 64-channel block (128-byte swizzle) + one
// 32-channel block (64-byte swizzle).  Roles: warp 0 TMA producer, warp 1 MMA issuer (18 tcgen05.mma of
// 128 x 192 x 16 and 18 of 128 x 96 x 16 per tile), warp 2 TMEM allocation, warps 4-15 epilogue (each warpgroup a
// third of the channels).
#include "../../include/msw_b200.h"
#include "msw_error.h"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace msw {

namespace cv {
constexpr int C = 96, HW_W = 16, TILE_PX = 128, ROWS_IN = 10, PX_IN = ROWS_IN * HW_W;   // 160 staged pixels
constexpr int THREADS = 512, STAGES = 2, TMEM_COLS = 512;       // 4 control warps + 12 epilogue warps
constexpr unsigned W0_TAP = C * 128u, W1_TAP = C * 64u;                 // bytes per tap: 64-ch block, 32-ch block
constexpr unsigned W0_BYTES = 9 * W0_TAP, W1_BYTES = 9 * W1_TAP;
constexpr unsigned A0_BYTES = PX_IN * 128u, A1_BYTES = PX_IN * 64u, A_STAGE = A0_BYTES + A1_BYTES;
constexpr unsigned OFF_W0 = 0, OFF_W1 = OFF_W0 + W0_BYTES, OFF_A = OFF_W1 + W1_BYTES;
constexpr unsigned OFF_BAR = OFF_A + STAGES * A_STAGE;                  // full[2], empty[2], tfull, d2_empty, wfull, d01_empty[2]
constexpr unsigned OFF_TMEM = OFF_BAR + 9 * 8u;
constexpr unsigned SMEM_BYTES = OFF_TMEM + 16u + 1024u;                 // + slack to align the base to 1024 B
}  // namespace cv

__device__ __forceinline__ unsigned cv_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cv_bar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void cv_bar_expect(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cv_bar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cv_bar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cv_tma_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void cv_tma_4d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// K-major operand descriptors (sm_100 version bit, LBO unused = 1): 128-byte swizzle = rows 128 B apart, 8-row
// groups 1024 B apart, layout type 2; 64-byte swizzle = rows 64 B apart, groups 512 B apart, layout type 4.
__device__ __forceinline__ uint64_t cv_desc128(unsigned addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t cv_desc64(unsigned addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void cv_mma(unsigned d_tmem, uint64_t a, uint64_t b, unsigned idesc, unsigned accumulate)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cv_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cv_ld16(unsigned taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(cv::THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                  __half *__restrict__ out, long long tiles, int dbg)
{
    using namespace cv;
    extern __shared__ unsigned char smem_dyn[];
    const unsigned base = (cv_smem(smem_dyn) + 1023u) & ~1023u;
    unsigned char *gen = smem_dyn + (base - cv_smem(smem_dyn));
    const unsigned bars = base + OFF_BAR;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (2 + s); };
    const unsigned tfull = bars + 8u * 4, d2_empty = bars + 8u * 5, wfull = bars + 8u * 6;
    auto d01_empty = [&](unsigned b) { return bars + 8u * (7 + b); };
    volatile uint32_t *s_tmem = reinterpret_cast<volatile uint32_t *>(gen + OFF_TMEM);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { cv_bar_init(full(s), 1); cv_bar_init(empty(s), 1); }
        cv_bar_init(tfull, 1);
        cv_bar_init(d2_empty, 12);
        cv_bar_init(d01_empty(0), 12);
        cv_bar_init(d01_empty(1), 12);
        cv_bar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(base + OFF_TMEM), "r"((unsigned)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *s_tmem;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer: the nine weight taps once, then one [10 rows x 16 px x 96 ch] block per tile
        cv_bar_expect(wfull, W0_BYTES + W1_BYTES);
        for (int t = 0; t < 9; ++t) {
            cv_tma_2d(base + OFF_W0 + t * W0_TAP, &map_w0, 0, t * C, wfull);
            cv_tma_2d(base + OFF_W1 + t * W1_TAP, &map_w1, 64, t * C, wfull);
        }
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const unsigned s = it % STAGES, ph = (it / STAGES) & 1u;
            const int n = (int)(tile >> 1), y0 = (int)(tile & 1) * 8;
            cv_bar_wait(empty(s), ph ^ 1u);
            cv_bar_expect(full(s), A_STAGE);
            cv_tma_4d(base + OFF_A + s * A_STAGE, &map_a0, 0, 0, y0 - 1, n, full(s));
            cv_tma_4d(base + OFF_A + s * A_STAGE + A0_BYTES, &map_a1, 64, 0, y0 - 1, n, full(s));
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer.  idesc: D = F32, A = B = F16, K-major, M = 128, N = 192 or 96.
        // One MMA covers the taps dx = -1 and dx = 0 together (their weight blocks are adjacent in shared memory
        // and their accumulators adjacent in TMEM, N = 192), a second one dx = +1 (N = 96): with three N = 96
        // MMAs the A operand is read from shared memory three times and the tensor pipe waits for it (7 KB of
        // operands per 48-cycle MMA > 128 B/clk).  TMEM: D_-1|D_0 double buffered at columns 0 / 192, D_+1 single
        // buffered at 384 (2 x 288 columns do not fit).  The epilogue copies its D_+1 values to registers first
        // and releases that buffer at once, so the N = 96 MMAs of the next tile wait for one TMEM load, not for
        // the whole epilogue; the N = 192 MMAs only need the epilogue of the tile before the previous one.
        constexpr unsigned idesc192 = (1u << 4) | ((unsigned)(2 * C >> 3) << 17) | ((unsigned)(TILE_PX >> 4) << 24);
        constexpr unsigned idesc96 = (1u << 4) | ((unsigned)(C >> 3) << 17) | ((unsigned)(TILE_PX >> 4) << 24);
        cv_bar_wait(wfull, 0);
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const unsigned s = it % STAGES, ph = (it / STAGES) & 1u;
            cv_bar_wait(full(s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned a0 = base + OFF_A + s * A_STAGE, a1 = a0 + A0_BYTES;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                // part 0 reuses the D_-1|D_0 buffer of tile it-2, part 1 the D_+1 buffer of tile it-1
                cv_bar_wait(part == 0 ? d01_empty(it & 1u) : d2_empty, part == 0 ? ((it >> 1) & 1u) ^ 1u : (it & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d = part == 0 ? tmem + (it & 1u) * 2u * C : tmem + 4u * C;
                const unsigned idesc = part == 0 ? idesc192 : idesc96;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int tap = dy * 3 + 2 * part;                       // first weight block of this MMA
#pragma unroll
                    for (int k = 0; k < 4; ++k)                              // channels 0..63: four 16-channel steps
                        cv_mma(d, cv_desc128(a0 + dy * HW_W * 128u) + 2u * k, cv_desc128(base + OFF_W0 + tap * W0_TAP) + 2u * k,
                               idesc, (dy | k) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 2; ++k)                              // channels 64..95
                        cv_mma(d, cv_desc64(a1 + dy * HW_W * 64u) + 2u * k, cv_desc64(base + OFF_W1 + tap * W1_TAP) + 2u * k,
                               idesc, 1u);
                }
            }
            cv_commit(empty(s));
            cv_commit(tfull);
        }
    } else if (warp >= 4) {
        // ---- epilogue: thread = pixel (TMEM lane q*32 + lane), each of the three warpgroups a third of the channels
        const int q = warp & 3, third = (warp - 4) >> 2;
        const int x = lane & 15;
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            cv_bar_wait(tfull, it & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned trow = tmem + ((unsigned)(q * 32) << 16) + third * 32;
            const unsigned d01 = trow + (it & 1u) * 2u * C, d2 = trow + 4u * C;
            __half *dst = out + ((tile >> 1) * 256 + (tile & 1) * 128 + q * 32 + lane) * (long long)C + third * 32;
            uint32_t vp[32];                       // D_+1 (contributes to the pixel on its left): copy out, release
            {
                uint32_t lo[16], hi[16];
                cv_ld16(d2, lo);
                cv_ld16(d2 + 16, hi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) { vp[j] = lo[j]; vp[16 + j] = hi[j]; }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) cv_bar_arrive(d2_empty);
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                if (dbg & 1) break;
                uint32_t vm[16], v0[16];
                cv_ld16(d01 + c0, vm);             // D_-1: contributes to the pixel on its right
                cv_ld16(d01 + C + c0, v0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float r[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(vm[j + e]), 1);
                        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(vp[c0 + j + e]), 1);
                        r[e] = __uint_as_float(v0[j + e]) + (x > 0 ? left : 0.0f) + (x < 15 ? right : 0.0f);
                    }
                    const __half2 h = __floats2half2_rn(r[0], r[1]);
                    packed[j >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                }
                // one 256-bit store per lane: the lanes of a warp write to 32 different 128-byte lines either way
                // (pixels are 192 B apart), so the LSU cost is per instruction, not per byte
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             :: "l"(dst + c0), "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]),
                                "r"(packed[4]), "r"(packed[5]), "r"(packed[6]), "r"(packed[7]) : "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) cv_bar_arrive(d01_empty(it & 1u));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((unsigned)TMEM_COLS) : "memory");
    }
}

typedef CUresult (*CvEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CvEncodeFn cv_encode_fn()
{
    static const CvEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<CvEncodeFn>(p);
    }();
    return fn;
}

}  // namespace msw

extern "C" int msw_conv3x3(const void *x16, const void *w_taps16, void *y16, int64_t n, int32_t H, int32_t W,
                           int32_t C, void *stream)
{
    using namespace msw;
    if (!x16 || !w_taps16 || !y16) return fail(MSW_ERR_NULL, "msw_conv3x3: NULL pointer");
    if (H != 16 || W != 16 || C != cv::C)
        return fail(MSW_ERR_BAD_SHAPE, "msw_conv3x3: only 16x16 boards with 96 channels (got %dx%d, C=%d)", H, W, C);
    if (n < 0 || n > 0x3fffffffLL) return fail(MSW_ERR_BAD_SHAPE, "msw_conv3x3: n=%lld", (long long)n);
    if ((((uintptr_t)x16 | (uintptr_t)w_taps16 | (uintptr_t)y16) & 15u) != 0)
        return fail(MSW_ERR_ALIGN, "msw_conv3x3: tensors must be 16-byte aligned");
    if (n == 0) return MSW_OK;
    if (!cv_encode_fn()) return fail(MSW_ERR_ARG, "msw_conv3x3: cuTensorMapEncodeTiled is not available");
    CUtensorMap ma0, ma1, mw0, mw1;
    {
        // activation [n][16][16][96] fp16: dims innermost first
        const cuuint64_t dims[4] = {(cuuint64_t)C, 16, 16, (cuuint64_t)n};
        const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 16, (cuuint64_t)C * 2 * 256};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const cuuint32_t box0[4] = {64, 16, cv::ROWS_IN, 1}, box1[4] = {32, 16, cv::ROWS_IN, 1};
        const CUresult r0 = cv_encode_fn()(&ma0, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(x16), dims, strides,
                                           box0, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const CUresult r1 = cv_encode_fn()(&ma1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(x16), dims, strides,
                                           box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r0 != CUDA_SUCCESS || r1 != CUDA_SUCCESS)
            return fail(MSW_ERR_ARG, "msw_conv3x3: activation tensor map failed (%d, %d)", (int)r0, (int)r1);
    }
    {
        // weights [9 taps * 96 co][96 ci] fp16
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)9 * C};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        const cuuint32_t estr[2] = {1, 1};
        const cuuint32_t box0[2] = {64, (cuuint32_t)C}, box1[2] = {32, (cuuint32_t)C};
        const CUresult r0 = cv_encode_fn()(&mw0, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(w_taps16), dims, strides,
                                           box0, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const CUresult r1 = cv_encode_fn()(&mw1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(w_taps16), dims, strides,
                                           box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r0 != CUDA_SUCCESS || r1 != CUDA_SUCCESS)
            return fail(MSW_ERR_ARG, "msw_conv3x3: weight tensor map failed (%d, %d)", (int)r0, (int)r1);
    }
    static thread_local bool configured = false;
    if (!configured) {
        MSW_CUDA_TRY(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cv::SMEM_BYTES));
        configured = true;
    }
    int dev = 0, sms = 0;
    MSW_CUDA_TRY(cudaGetDevice(&dev));
    MSW_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = 2 * (long long)n;
    const long long grid = tiles < sms ? tiles : sms;
    const char *e = getenv("MSW_CONV_DBG");
    conv3x3_tc_kernel<<<(unsigned)grid, cv::THREADS, cv::SMEM_BYTES, (cudaStream_t)stream>>>(ma0, ma1, mw0, mw1,
                                                                                            (__half *)y16, tiles, e ? atoi(e) : 0);
    MSW_CUDA_TRY(cudaGetLastError());
    return MSW_OK;
}
