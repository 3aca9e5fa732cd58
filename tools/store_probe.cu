// store_probe.cu -- what is the write-only HBM roofline of this B200, and which store flavour
// reaches it?  Development tool (not part of the product): the env step kernel is a pure
// streaming-write kernel, while MEASURED_PEAKS.json's hbm_gbs is a read+write copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/store_probe tools/store_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

enum { ST_DEFAULT = 0, ST_CS = 1, ST_NOALLOC = 2 };

template <int KIND>
__device__ __forceinline__ void store16(float4 *p, float4 v)
{
    if (KIND == ST_DEFAULT) *p = v;
    else if (KIND == ST_CS) __stcs(p, v);
    else asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// each warp owns consecutive CHUNK-byte chunks (like one board's 10 KB of obs)
template <int KIND>
__global__ void __launch_bounds__(256) fill_chunks(float4 *out, long long n_chunks, int f4_per_chunk)
{
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const long long total = (long long)gridDim.x * wpb;
    const float4 v = make_float4(0.f, 1.f, 0.f, 0.f);
    for (long long c = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); c < n_chunks; c += total) {
        float4 *base = out + c * f4_per_chunk;
#pragma unroll 4
        for (int i = lane; i < f4_per_chunk; i += 32) store16<KIND>(base + i, v);
    }
}

__global__ void __launch_bounds__(256) fill_v8(float *out, long long n_chunks, int f_per_chunk)
{
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const long long total = (long long)gridDim.x * wpb;
    for (long long c = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); c < n_chunks; c += total) {
        float *base = out + c * f_per_chunk;
#pragma unroll 4
        for (int i = lane * 8; i < f_per_chunk; i += 256)
            asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(base + i), "f"(1.0f) : "memory");
    }
}

// flat grid-stride fill (every warp instruction = 512 contiguous bytes, neighbouring warps adjacent)
template <int KIND>
__global__ void __launch_bounds__(256) fill_flat(float4 *out, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4 v = make_float4(0.f, 1.f, 0.f, 0.f);
#pragma unroll 4
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) store16<KIND>(out + i, v);
}

// TMA bulk store: one CTA stages a tile in shared memory once, then streams it out with
// cp.async.bulk.global.shared::cta (UBLKCP), up to 8 bulk copies in flight per CTA.
__global__ void __launch_bounds__(128) fill_bulk(char *out, long long n_tiles, int tile_bytes)
{
    extern __shared__ __align__(128) char smem[];
    for (int i = threadIdx.x * 16; i < tile_bytes; i += blockDim.x * 16) *reinterpret_cast<float4 *>(smem + i) = make_float4(0.f, 1.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + t * tile_bytes), "r"(s), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) copy_flat(const float4 *__restrict__ in, float4 *__restrict__ out, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 4
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

template <typename F>
static void timeit(const char *name, double bytes, F launch)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0;
    const int reps = 20;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best; sum += ms;
    }
    CK(cudaGetLastError());
    printf("%-44s best %8.1f GB/s   mean %8.1f GB/s   (%.1f us)\n", name, bytes / best / 1e6, bytes / (sum / reps) / 1e6, best * 1e3);
}

int main()
{
    const long long n_chunks = 65536;
    const int chunk_bytes = 10496;                       // one 16x16 board: 10 planes * 1 KiB + 256 B mask
    const long long bytes = n_chunks * chunk_bytes;      // 688 MB
    char *buf, *src;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&src, bytes));
    CK(cudaMemset(src, 1, bytes));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d, buffer %.0f MB\n", sms, bytes / 1e6);
    for (int mult : {4, 8, 16}) {
        const int grid = sms * mult;
        char nm[96];
        snprintf(nm, 96, "chunks st.default grid=%dxSM", mult); timeit(nm, (double)bytes, [&] { fill_chunks<ST_DEFAULT><<<grid, 256>>>((float4 *)buf, n_chunks, chunk_bytes / 16); });
        snprintf(nm, 96, "chunks st.cs      grid=%dxSM", mult); timeit(nm, (double)bytes, [&] { fill_chunks<ST_CS><<<grid, 256>>>((float4 *)buf, n_chunks, chunk_bytes / 16); });
        snprintf(nm, 96, "chunks st.noalloc grid=%dxSM", mult); timeit(nm, (double)bytes, [&] { fill_chunks<ST_NOALLOC><<<grid, 256>>>((float4 *)buf, n_chunks, chunk_bytes / 16); });
        snprintf(nm, 96, "chunks st.v8      grid=%dxSM", mult); timeit(nm, (double)bytes, [&] { fill_v8<<<grid, 256>>>((float *)buf, n_chunks, chunk_bytes / 4); });
        snprintf(nm, 96, "flat   st.default grid=%dxSM", mult); timeit(nm, (double)bytes, [&] { fill_flat<ST_DEFAULT><<<grid, 256>>>((float4 *)buf, bytes / 16); });
        snprintf(nm, 96, "flat   st.cs      grid=%dxSM", mult); timeit(nm, (double)bytes, [&] { fill_flat<ST_CS><<<grid, 256>>>((float4 *)buf, bytes / 16); });
    }
    CK(cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int tile : {10240, 20480, 40960}) {
        for (int mult : {1, 2, 4}) {
            char nm[96];
            snprintf(nm, 96, "TMA bulk store tile=%dB grid=%dxSM", tile, mult);
            const long long tiles = bytes / tile;
            timeit(nm, (double)tiles * tile, [&] { fill_bulk<<<sms * mult, 128, tile>>>(buf, tiles, tile); });
        }
    }
    timeit("cudaMemsetAsync", (double)bytes, [&] { CK(cudaMemsetAsync(buf, 0, bytes)); });
    timeit("copy kernel (read+write bytes)", 2.0 * bytes, [&] { copy_flat<<<sms * 8, 256>>>((const float4 *)src, (float4 *)buf, bytes / 16); });
    timeit("cudaMemcpyAsync D2D (read+write bytes)", 2.0 * bytes, [&] { CK(cudaMemcpyAsync(buf, src, bytes, cudaMemcpyDeviceToDevice)); });
    return 0;
}
