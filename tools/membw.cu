// membw.cu -- what the B200's HBM gives a plain streaming kernel (the ceilings the path's kernels are
// compared with in profiles/README.md): read-only and write-only streams, 128-bit and 256-bit accesses,
// at the 160 MB size of a cfg2 product (rotating over 16 buffers so nothing is L2-resident) and at 4 GB.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/membw tools/membw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct alignas(32) u256 { uint64_t a, b, c, d; };

__device__ __forceinline__ uint4 ld128(const uint4 *p) { return __ldcs(p); }
__device__ __forceinline__ u256 ld256(const u256 *p) {
    u256 v;
    asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ void st256(u256 *p, const u256 &v) {
    asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}

template <int U>
__global__ void __launch_bounds__(256) read128(const uint4 *__restrict__ p, size_t n, uint64_t *sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    uint32_t acc = 0;
    for (; i + (size_t)(U - 1) * blockDim.x < n; i += stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld128(p + i + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int U>
__global__ void __launch_bounds__(256) read256(const u256 *__restrict__ p, size_t n, uint64_t *sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    uint64_t acc = 0;
    for (; i + (size_t)(U - 1) * blockDim.x < n; i += stride) {
        u256 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld256(p + i + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].a ^ v[u].b ^ v[u].c ^ v[u].d;
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int U>
__global__ void __launch_bounds__(256) write128(uint4 *__restrict__ p, size_t n, uint32_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    const uint4 v = make_uint4(seed, threadIdx.x, blockIdx.x, 7u);
    for (; i + (size_t)(U - 1) * blockDim.x < n; i += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(p + i + (size_t)u * blockDim.x, v);
    }
}
template <int U>
__global__ void __launch_bounds__(256) write256(u256 *__restrict__ p, size_t n, uint32_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    u256 v;
    v.a = seed; v.b = threadIdx.x; v.c = blockIdx.x; v.d = 7;
    for (; i + (size_t)(U - 1) * blockDim.x < n; i += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) st256(p + i + (size_t)u * blockDim.x, v);
    }
}

template <typename F>
float time_us(F launch, int nbuf, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int b = 0; b < nbuf; ++b) launch(b);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) for (int b = 0; b < nbuf; ++b) launch(b);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / (reps * nbuf);
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint64_t *sink; CK(cudaMalloc(&sink, 8));
    const size_t sizes[2] = {160000000ull, 4000000000ull};
    for (int si = 0; si < 2; ++si) {
        const size_t bytes = sizes[si];
        const int nbuf = si == 0 ? 16 : 2, reps = si == 0 ? 5 : 3;
        std::vector<void *> bufs(nbuf);
        for (auto &b : bufs) { CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 1, bytes)); }
        for (int ctas_per_sm : {2, 4, 8, 16}) {
            const int grid = sms * ctas_per_sm;
            float t;
            t = time_us([&](int b) { read128<4><<<grid, 256>>>((const uint4 *)bufs[b], bytes / 16, sink); }, nbuf, reps);
            printf("%5.2f GB read128 U=4  grid=%4d  %8.2f us  %7.1f GB/s\n", bytes / 1e9, grid, t, bytes / t / 1e3);
            t = time_us([&](int b) { read128<8><<<grid, 256>>>((const uint4 *)bufs[b], bytes / 16, sink); }, nbuf, reps);
            printf("%5.2f GB read128 U=8  grid=%4d  %8.2f us  %7.1f GB/s\n", bytes / 1e9, grid, t, bytes / t / 1e3);
            t = time_us([&](int b) { read256<4><<<grid, 256>>>((const u256 *)bufs[b], bytes / 32, sink); }, nbuf, reps);
            printf("%5.2f GB read256 U=4  grid=%4d  %8.2f us  %7.1f GB/s\n", bytes / 1e9, grid, t, bytes / t / 1e3);
            t = time_us([&](int b) { write128<4><<<grid, 256>>>((uint4 *)bufs[b], bytes / 16, 3u); }, nbuf, reps);
            printf("%5.2f GB write128 U=4 grid=%4d  %8.2f us  %7.1f GB/s\n", bytes / 1e9, grid, t, bytes / t / 1e3);
            t = time_us([&](int b) { write256<4><<<grid, 256>>>((u256 *)bufs[b], bytes / 32, 3u); }, nbuf, reps);
            printf("%5.2f GB write256 U=4 grid=%4d  %8.2f us  %7.1f GB/s\n", bytes / 1e9, grid, t, bytes / t / 1e3);
        }
        // two streams: overlapping tail and ramp of consecutive launches
        cudaStream_t s[2]; cudaStreamCreate(&s[0]); cudaStreamCreate(&s[1]);
        float t = time_us([&](int b) { read128<8><<<sms * 4, 256, 0, s[b & 1]>>>((const uint4 *)bufs[b], bytes / 16, sink); }, nbuf, reps);
        printf("%5.2f GB read128 U=8  2 streams (event span incl. only stream 0)  %8.2f us\n", bytes / 1e9, t);
        cudaDeviceSynchronize();
        for (auto &b : bufs) cudaFree(b);
    }
    return 0;
}
