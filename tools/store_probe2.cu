// store_probe2.cu -- follow-up to store_probe.cu: how do occupancy (resident CTAs/SM), grid size
// (static persistent vs dynamically scheduled CTAs) and L2 eviction hints change the achievable
// write-only bandwidth for the env kernel's pattern (each warp streams one contiguous 10,496 B chunk)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

enum { ST_CS = 0, ST_EVICT_FIRST = 1, ST_EVICT_LAST = 2, ST_WT = 3 };

template <int KIND>
__device__ __forceinline__ void store16(float4 *p, float4 v, uint64_t pol)
{
    if (KIND == ST_CS) __stcs(p, v);
    else if (KIND == ST_WT) __stwt(p, v);
    else asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

template <int KIND>
__global__ void fill_chunks(float4 *out, long long n_chunks, int f4_per_chunk)
{
    extern __shared__ char dyn[];      // only used to limit resident CTAs per SM
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const long long total = (long long)gridDim.x * wpb;
    uint64_t pol = 0;
    if (KIND == ST_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (KIND == ST_EVICT_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const float4 v = make_float4(0.f, 1.f, 0.f, 0.f);
    for (long long c = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); c < n_chunks; c += total) {
        float4 *base = out + c * f4_per_chunk;
#pragma unroll 4
        for (int i = lane; i < f4_per_chunk; i += 32) store16<KIND>(base + i, v, pol);
    }
}

// block-contiguous assignment: CTA j owns chunks [j*per, (j+1)*per): its 8 warps write neighbours
template <int KIND>
__global__ void fill_chunks_blocked(float4 *out, long long n_chunks, int f4_per_chunk, int per_block)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const float4 v = make_float4(0.f, 1.f, 0.f, 0.f);
    const long long c0 = (long long)blockIdx.x * per_block;
    for (int k = warp; k < per_block; k += wpb) {
        const long long c = c0 + k;
        if (c >= n_chunks) break;
        float4 *base = out + c * f4_per_chunk;
#pragma unroll 4
        for (int i = lane; i < f4_per_chunk; i += 32) store16<KIND>(base + i, v, 0);
    }
}

template <typename F>
static float timeit(F launch)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float sum = 0;
    const int reps = 20;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); sum += ms;
    }
    CK(cudaGetLastError());
    return sum / reps;
}

int main()
{
    const long long n_chunks = 65536;
    const int chunk_bytes = 10496;
    const long long bytes = n_chunks * chunk_bytes;
    char *buf;
    CK(cudaMalloc(&buf, bytes * 4));       // 4 slots like the bench ring
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(fill_chunks<ST_CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    printf("== chunks, st.cs: resident CTAs/SM (256 thr) x grid multiplier -> mean GB/s\n");
    for (int res : {2, 3, 4, 6, 8}) {
        const int smem = res >= 8 ? 0 : (227 * 1024 / res - 2048);
        printf("resident %d/SM:", res);
        for (int mult : {2, 3, 4, 6, 8, 12, 16, 24, 32, 55}) {
            if (mult < res) { printf("      -  "); continue; }
            int slot = 0;
            float ms = timeit([&] { fill_chunks<ST_CS><<<sms * mult, 256, smem>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16); });
            printf("  x%-2d %5.0f", mult, bytes / ms / 1e6);
        }
        printf("\n");
    }
    printf("== chunks, 8 resident, grid x8 / x16, store kinds\n");
    {
        int slot = 0;
        for (int mult : {8, 16}) {
            float a = timeit([&] { fill_chunks<ST_CS><<<sms * mult, 256>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16); });
            float b = timeit([&] { fill_chunks<ST_EVICT_FIRST><<<sms * mult, 256>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16); });
            float c = timeit([&] { fill_chunks<ST_EVICT_LAST><<<sms * mult, 256>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16); });
            float d = timeit([&] { fill_chunks<ST_WT><<<sms * mult, 256>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16); });
            printf("x%-2d  cs %5.0f  L2::evict_first %5.0f  L2::evict_last %5.0f  wt %5.0f GB/s\n", mult, bytes / a / 1e6, bytes / b / 1e6, bytes / c / 1e6, bytes / d / 1e6);
        }
    }
    printf("== block-contiguous chunks (CTA owns `per` consecutive boards), st.cs\n");
    for (int per : {8, 16, 32, 64, 128}) {
        int slot = 0;
        const int grid = (int)((n_chunks + per - 1) / per);
        float ms = timeit([&] { fill_chunks_blocked<ST_CS><<<grid, 256>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16, per); });
        printf("per=%-3d grid=%-5d %5.0f GB/s\n", per, grid, bytes / ms / 1e6);
    }
    printf("== block size (grid = 16 x SM x 256/threads)\n");
    for (int thr : {128, 256, 512, 1024}) {
        int slot = 0;
        float ms = timeit([&] { fill_chunks<ST_CS><<<sms * 16 * 256 / thr, thr>>>((float4 *)(buf + (slot++ & 3) * bytes), n_chunks, chunk_bytes / 16); });
        printf("threads=%-4d %5.0f GB/s\n", thr, bytes / ms / 1e6);
    }
    float ms = timeit([&] { CK(cudaMemsetAsync(buf, 0, bytes)); });
    printf("cudaMemsetAsync %5.0f GB/s\n", bytes / ms / 1e6);
    return 0;
}
