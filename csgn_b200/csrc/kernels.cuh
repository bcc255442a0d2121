// kernels.cuh -- internal launch interface between the C ABI (capi.cu) and the
// sm_100a kernels.  Device pointers + sizes + stream; every launcher returns the
// cudaError_t of the launch and bumps the library's launch counter.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "peer.cuh"

namespace csgn {

struct DeviceProps {
    int device = -1;
    int sm_count = 148;
    int cc = 100;
    size_t smem_optin = 227 * 1024;
};
const DeviceProps &device_props();
void set_device_props(const DeviceProps &p);
void count_launch(unsigned n = 1);
uint64_t launches();
// true when the library was built with -DCSGN_BUILD_VARIANTS (the losing kernel variants kept for A/B sweeps)
bool build_has_variants();

// Integer environment knob (tuning sweeps only); `dflt` when unset or malformed.
long env_long(const char *name, long dflt);

// Fused multiply -> fold: what launch_mul needs to also count the satisfied blocks of the product.
struct MulFold {
    const uint64_t *mask;        // key mask, L words, device
    const uint64_t *host_mask;   // the same on the host (optional): small masks ride in the kernel parameters
    uint64_t *scratch;           // the launch's fold scratch word (zero between launches)
    uint64_t *count_out;         // device word for the total; may be null when `peer` is given
    const PeerPush *peer;        // optional: sharded decrypt (push / publish / collect in the same kernel)
    bool overlapped = false;     // the caller runs this launch next to other kernels (batch lanes, automatic lanes)
};
// K1  out[(i*T2+j)*L+k] = a[i*L+k] & b[j*L+k]        (reference src/Ciphertext.cpp:153-163)
// With `fold` the same launch also decrypt-folds the product (src/SecretKey.cpp:131-140) while its units are in
// registers; `out` may then be null (count only, nothing is stored).  cudaErrorNotSupported when a shape has no
// fused kernel (mul_fold_supported(L) is false): the caller multiplies and folds in two launches.
cudaError_t launch_mul(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2, uint32_t L,
                       uint64_t *out, cudaStream_t stream, const MulFold *fold = nullptr);
bool mul_fold_supported(uint32_t L);

// K2  out = a || b                                     (reference src/Ciphertext.cpp:107-122)
// Either source may be null/empty; out may alias a (append in place) when a == out.
cudaError_t launch_concat(const uint64_t *a, uint64_t n_words_a, const uint64_t *b, uint64_t n_words_b,
                          uint64_t *out, cudaStream_t stream);

// out[0] = sum of in[0..n): the partial counts of a ciphertext held as several segments (a lazy sum)
cudaError_t launch_sum_words(const uint64_t *in, uint32_t n, uint64_t *out, cudaStream_t stream);

// K3  count of blocks with all_w((v[w] & M[w]) == M[w]) (reference src/SecretKey.cpp:126-140)
// `scratch` is one zero-initialised uint64 (count | CTA ticket << 40, fold.cuh) that the kernel
// leaves zeroed again; the total is written to *count_out (device memory).  `overlapped`: the caller runs
// this fold next to other kernels (batch lanes), so several shorter waves of CTAs back-fill better than one.
// `host_mask` (optional) is a host copy of the same L words: small masks ride in the
// kernel parameters instead of being fetched from global memory.
cudaError_t launch_decrypt_count(const uint64_t *v, uint64_t T, uint32_t L, const uint64_t *mask,
                                 const uint64_t *host_mask, uint64_t *scratch, uint64_t *count_out,
                                 cudaStream_t stream, const PeerPush *peer = nullptr, bool overlapped = false);
// `peer` (optional, sharded decrypt): the kernel's last CTA also pushes the count into every
// rank's mailbox and, when peer->collect_n > 0, collects the batch's totals (peer.cuh);
// count_out may then be null.  launch_peer_exchange does the push/collect without a fold.
cudaError_t launch_peer_exchange(const PeerPush &pp, bool do_push, uint64_t value, uint64_t *count_out,
                                 cudaStream_t stream);

// K4  out_bit[i] = in_bit[perm[i]] for every block     (reference src/Ciphertext.cpp:24-69)
// src_map[i] = (perm[i]>>6)<<6 | (63 - (perm[i]&63)): source word and right-shift.
// slice_map (optional, 64*L entries; see permute.cu) enables the bit-sliced tile kernel.
// plane_map (optional, 64*L entries): the same gather for the plane kernel (long blocks; slices kept in the tile's own
// 32 x 2L array): byte offset (j_src*2L + c_src)*4, or 4*32*2L (zero words) for pad bits.
cudaError_t launch_permute(const uint64_t *in, uint64_t T, uint32_t L, uint32_t N, const uint32_t *src_map,
                           const uint32_t *slice_map, const uint32_t *plane_map, uint64_t *out, cudaStream_t stream);
bool permute_sliced_supported(uint32_t L);
bool permute_plane_supported(uint32_t L);
// Words between consecutive slice rows of a tile in shared memory (32 slices + padding; 16-byte
// aligned rows, conflict-free 128-bit stores).  The slice map's byte offsets are built with it.
constexpr uint32_t kPermSliceStride = 36;

// n fresh encryptions (one block per plaintext bit) with Philox-4x32-10 keyed by `seed`; see encrypt.cu.
cudaError_t launch_encrypt_batch(const uint8_t *bits, uint64_t n, uint64_t first_block, uint32_t L, uint64_t pad_mask,
                                 const uint64_t *mask, const uint64_t *positions, uint32_t D, uint64_t seed,
                                 uint64_t *out, cudaStream_t stream);

// xor / wrapping sum / sum(w[i]*(2i+1)) of n_words words, accumulated into acc[0..2]
// (device, must be zeroed by the caller).
cudaError_t launch_checksum(const uint64_t *v, uint64_t n_words, uint64_t *acc, cudaStream_t stream);

}  // namespace csgn
