// misc.cu -- K2 (ciphertext add = block concatenation) and the device checksum.
//
// K2: out = a || b            (reference src/Ciphertext.cpp:107-122, :215-223)
// Roofline: HBM copy bandwidth, 16*L bytes per block moved (8*L read + 8*L written).
// operator+= appends in place (out == a): only b moves.
#include "kernels.cuh"
#include "launch.cuh"

#include <algorithm>

namespace csgn {
namespace {

constexpr int kCopyThreads = 256;
constexpr int kCopyUnroll = 4;

// Two-source streaming copy.  VecT = uint4 when every pointer and both lengths are
// 16-byte granular, else uint64_t.
template <typename VecT>
__global__ void __launch_bounds__(kCopyThreads)
concat_kernel(const VecT *__restrict__ a, const uint64_t na, const VecT *__restrict__ b, const uint64_t nb,
              VecT *__restrict__ out) {
    const uint64_t total = na + nb;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * kCopyUnroll;
    pdl_enter();
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * kCopyUnroll + threadIdx.x; base < total; base += stride) {
        VecT v[kCopyUnroll];
        bool live[kCopyUnroll];
#pragma unroll
        for (int u = 0; u < kCopyUnroll; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            live[u] = i < total;
            if (live[u]) v[u] = (i < na) ? __ldcs(a + i) : __ldcs(b + (i - na));
        }
#pragma unroll
        for (int u = 0; u < kCopyUnroll; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (live[u]) __stcs(out + i, v[u]);
        }
    }
}

__global__ void __launch_bounds__(256)
checksum_kernel(const uint64_t *__restrict__ v, const uint64_t n_words, uint64_t *acc) {
    uint64_t x = 0, s = 0, h = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    pdl_enter();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint64_t w = __ldcs(v + i);
        x ^= w;
        s += w;
        h += w * (2ull * i + 1ull);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, off);
        s += __shfl_xor_sync(0xffffffffu, s, off);
        h += __shfl_xor_sync(0xffffffffu, h, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicXor(reinterpret_cast<unsigned long long *>(acc), (unsigned long long)x);
        atomicAdd(reinterpret_cast<unsigned long long *>(acc + 1), (unsigned long long)s);
        atomicAdd(reinterpret_cast<unsigned long long *>(acc + 2), (unsigned long long)h);
    }
}

// out[0] = in[0] + ... + in[n-1]   (the partial counts of a ciphertext held as several segments)
__global__ void __launch_bounds__(32)
sum_words_kernel(const uint64_t *__restrict__ in, const uint32_t n, uint64_t *out) {
    pdl_enter();
    uint64_t s = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 32) s += in[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0) *out = s;
}

}  // namespace

cudaError_t launch_sum_words(const uint64_t *in, uint32_t n, uint64_t *out, cudaStream_t stream) {
    count_launch();
    return launch_kernel(sum_words_kernel, 1, 32, 0, stream, in, n, out);
}

cudaError_t launch_concat(const uint64_t *a, uint64_t n_words_a, const uint64_t *b, uint64_t n_words_b,
                          uint64_t *out, cudaStream_t stream) {
    if (a == out) {  // append in place: only b moves
        out += n_words_a;
        a = nullptr;
        n_words_a = 0;
    }
    if (n_words_a + n_words_b == 0) return cudaSuccess;
    const DeviceProps &dp = device_props();
    const bool vec = (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                        reinterpret_cast<uintptr_t>(out)) & 15u) == 0) &&
                     (n_words_a % 2 == 0) && (n_words_b % 2 == 0);
    const uint64_t units = vec ? (n_words_a + n_words_b) / 2 : (n_words_a + n_words_b);
    const uint64_t per_cta = (uint64_t)kCopyThreads * kCopyUnroll;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>((units + per_cta - 1) / per_cta, (uint64_t)dp.sm_count * 16));
    count_launch();
    if (vec)
        return launch_kernel(concat_kernel<uint4>, grid, kCopyThreads, 0, stream, reinterpret_cast<const uint4 *>(a),
                             n_words_a / 2, reinterpret_cast<const uint4 *>(b), n_words_b / 2,
                             reinterpret_cast<uint4 *>(out));
    return launch_kernel(concat_kernel<uint64_t>, grid, kCopyThreads, 0, stream, a, n_words_a, b, n_words_b, out);
}

cudaError_t launch_checksum(const uint64_t *v, uint64_t n_words, uint64_t *acc, cudaStream_t stream) {
    if (n_words == 0) return cudaSuccess;
    const DeviceProps &dp = device_props();
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>((n_words + 255) / 256, (uint64_t)dp.sm_count * 8));
    count_launch();
    return launch_kernel(checksum_kernel, grid, 256, 0, stream, v, n_words, acc);
}

}  // namespace csgn
