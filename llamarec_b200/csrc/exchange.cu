// Peer-memory exchange for the row-sharded, data-parallel retrieval step (one process per GPU, buffers
// mapped into every process of the node over NVLink / NVSwitch).
//
// The reference has no multi-device retrieval path; these kernels replace what a
// collective library would do around the scoring kernel:
//   lrb_peer_push   "all-gather by stores": every rank writes its users' exchange records (user state,
//                   sorted exclusion list, filter) straight into its slot of every peer's gather buffer.
//   (the matching "all-to-all by stores" is fused into the merge kernel: lrb_merge_metrics_scatter writes
//    each user's merged local list into the recv buffer of the rank that owns the user.)
// Ordering between ranks is the caller's (a signal-pad barrier after the stores; buffers are double
// buffered, see llamarec_b200/sharded.py).
#include "api_util.h"

#include <cstdint>

namespace lrb {
namespace xchg {

constexpr int MAX_ARRAYS = 4;
constexpr int MAX_DST = 16;

struct PushParams {
  const uint4* src[MAX_ARRAYS];
  uint4* dst[MAX_ARRAYS][MAX_DST];
  unsigned long long n16[MAX_ARRAYS];   // 16-byte units per array
  int n_arrays, n_dst;
};

// grid = (chunks, n_dst, n_arrays); every thread moves 16 bytes per iteration, coalesced on both sides.
__global__ void __launch_bounds__(256) peer_push_kernel(const PushParams p) {
  const int a = blockIdx.z, d = blockIdx.y;
  const uint4* __restrict__ src = p.src[a];
  uint4* __restrict__ dst = p.dst[a][d];
  const unsigned long long n = p.n16[a];
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

}  // namespace xchg
}  // namespace lrb

extern "C" int lrb_peer_push(const void* const* src_host, const size_t* bytes_host, int n_arrays,
                             void* const* dst_host, int n_dst, void* stream) {
  using namespace lrb;
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  LRB_REQUIRE(src_host && bytes_host && dst_host, "lrb_peer_push: null pointer");
  LRB_REQUIRE(n_arrays >= 1 && n_arrays <= xchg::MAX_ARRAYS, "lrb_peer_push: 1..%d arrays", xchg::MAX_ARRAYS);
  LRB_REQUIRE(n_dst >= 1 && n_dst <= xchg::MAX_DST, "lrb_peer_push: 1..%d destinations", xchg::MAX_DST);
  xchg::PushParams p = {};
  p.n_arrays = n_arrays;
  p.n_dst = n_dst;
  unsigned long long max_n = 0;
  for (int a = 0; a < n_arrays; ++a) {
    LRB_REQUIRE(src_host[a] != nullptr && bytes_host[a] % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(src_host[a]) % 16 == 0,
                "lrb_peer_push: array %d must be 16-byte aligned and a multiple of 16 bytes", a);
    p.src[a] = static_cast<const uint4*>(src_host[a]);
    p.n16[a] = bytes_host[a] / 16;
    if (p.n16[a] > max_n) max_n = p.n16[a];
    for (int d = 0; d < n_dst; ++d) {
      void* dst = dst_host[a * n_dst + d];
      LRB_REQUIRE(dst != nullptr && reinterpret_cast<uintptr_t>(dst) % 16 == 0,
                  "lrb_peer_push: destination %d of array %d is null or misaligned", d, a);
      p.dst[a][d] = static_cast<uint4*>(dst);
    }
  }
  if (max_n == 0) return LRB_OK;
  unsigned chunks = static_cast<unsigned>((max_n + 256 * 4 - 1) / (256 * 4));   // ~4 units per thread
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, static_cast<unsigned>(n_dst), static_cast<unsigned>(n_arrays));
  xchg::peer_push_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}
