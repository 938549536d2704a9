// Layout / preparation kernels: item-table conversion and per-user sequence bookkeeping.
#include "api_util.h"
#include "common.cuh"

#include <climits>

namespace lrb {
namespace {

// ---- table: fp32 -> bf16 copy of a row range, bias padded with -inf ------------------------
__global__ void prepare_table_kernel(const float* __restrict__ table, const float* __restrict__ bias,
                                     long long row_begin, long long rows, __nv_bfloat16* __restrict__ out,
                                     float* __restrict__ bias_pad, uint4* __restrict__ bias_blk,
                                     long long rows_pad) {
  const long long n4 = rows * (LRB_D / 4);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const float4* src = reinterpret_cast<const float4*>(table + row_begin * LRB_D);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = pk;
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows_pad; i += stride) {
    const float bv = i < rows ? bias[row_begin + i] : -INFINITY;
    bias_pad[i] = bv;
    if (bias_blk != nullptr) {
      // folded-bias block consumed by the scoring GEMM as a K=16 slab: bias = hi + mid + lo (three
      // bf16 terms reproduce the fp32 value to ~2^-24), stored as a plain row-major [rows_pad][16] bf16
      // matrix (32 B per item: K columns 0..2 = hi, mid, lo, the rest zero); TMA brings it into shared
      // memory with the 32-byte swizzle the UMMA descriptor expects.
      const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
      float rest = i < rows ? bv - __bfloat162float(hi) : 0.f;
      const __nv_bfloat16 mid = __float2bfloat16_rn(rest);
      rest = i < rows ? rest - __bfloat162float(mid) : 0.f;
      const __nv_bfloat16 lo = __float2bfloat16_rn(rest);
      uint4 k_lo;
      k_lo.x = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) |
               (static_cast<uint32_t>(__bfloat16_as_ushort(mid)) << 16);
      k_lo.y = static_cast<uint32_t>(__bfloat16_as_ushort(lo));
      k_lo.z = 0u; k_lo.w = 0u;
      bias_blk[2 * i] = k_lo;                               // K columns 0..7
      bias_blk[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);     // K columns 8..15
    }
  }
}

// ---- sequences: first token, token count, sorted exclusion list + bloom --------------------
// One CTA per user.  L <= LRB_MAX_LEN.
constexpr int SEQ_THREADS = 128;

__global__ void __launch_bounds__(SEQ_THREADS)
prepare_sequences_kernel(const void* __restrict__ ids, int id_bytes, int B, int L, int all_positions,
                         int* __restrict__ tok_first, int* __restrict__ tok_offset,
                         int* __restrict__ excl_sorted, uint32_t* __restrict__ excl_bloom, int stride) {
  __shared__ int s_ids[LRB_MAX_LEN + 1];
  __shared__ int s_first;
  __shared__ uint32_t s_bloom[4];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) s_first = L;
  if (tid < 4) s_bloom[tid] = 0u;
  __syncthreads();
  for (int t = tid; t < L; t += SEQ_THREADS) {
    const int id = static_cast<int>(load_id(ids, static_cast<size_t>(b) * L + t, id_bytes));
    s_ids[t] = id;
    if (id > 0) atomicMin(&s_first, t);
  }
  if (tid == 0) s_ids[L] = 0;   // the padding item is always excluded (trainer/lru.py:38)
  __syncthreads();
  if (tid == 0) {
    int first = all_positions ? 0 : s_first;
    if (first >= L) first = L - 1;   // empty history: keep the last position so u is defined
    tok_first[b] = first;
    tok_offset[b + 1] = L - first;   // counts; turned into offsets by the scan kernel
    if (b == 0) tok_offset[0] = 0;
  }
  if (excl_sorted == nullptr) return;
  // rank sort of the L+1 ids (duplicates keep their relative order; duplicates are harmless
  // for the binary search that consumes the list)
  const int n = L + 1;
  int* out = excl_sorted + static_cast<size_t>(b) * stride;
  for (int t = tid; t < n; t += SEQ_THREADS) {
    const int v = s_ids[t];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const int w = s_ids[j];
      rank += (w < v) || (w == v && j < t);
    }
    out[rank] = v;
    atomicOr(&s_bloom[(v >> 5) & 3], 1u << (v & 31));
  }
  for (int t = n + tid; t < stride; t += SEQ_THREADS) out[t] = INT_MAX;
  __syncthreads();
  if (tid < 4) excl_bloom[static_cast<size_t>(b) * 4 + tid] = s_bloom[tid];
}

// single-CTA inclusive scan of tok_offset[1..B] (counts -> offsets)
__global__ void __launch_bounds__(1024) scan_counts_kernel(int* __restrict__ tok_offset, int B) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + tid;
    int v = i < B ? tok_offset[i + 1] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int prefix = s_carry + (warp > 0 ? s_warp[warp - 1] : 0);
    if (i < B) tok_offset[i + 1] = prefix + v;
    __syncthreads();
    if (tid == 1023) s_carry = prefix + v;
    __syncthreads();
  }
}

}  // namespace
}  // namespace lrb

extern "C" {

int lrb_excl_stride(int L) { return ((L + 1) + 3) & ~3; }

size_t lrb_bias_blk_bytes(int64_t rows) { return static_cast<size_t>((rows + 255) / 256) * 8192; }

int lrb_prepare_table(const float* table_f32, const float* bias_f32, int64_t row_begin, int64_t rows,
                      void* table_bf16, float* bias_pad, void* bias_blk, void* stream) {
  using namespace lrb;
  LRB_REQUIRE(table_f32 && bias_f32 && table_bf16 && bias_pad, "lrb_prepare_table: null pointer");
  LRB_REQUIRE(rows > 0 && row_begin >= 0, "lrb_prepare_table: bad row range");
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  const long long rows_pad = (rows + 255) / 256 * 256;
  int sms = device_sm_count();
  prepare_table_kernel<<<sms * 8, 256, 0, as_stream(stream)>>>(
      table_f32, bias_f32, row_begin, rows, static_cast<__nv_bfloat16*>(table_bf16), bias_pad,
      static_cast<uint4*>(bias_blk), rows_pad);
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}

int lrb_prepare_sequences(const void* ids, int id_bytes, int B, int L, int all_positions, int32_t* tok_first,
                          int32_t* tok_offset, int32_t* excl_sorted, uint32_t* excl_bloom, void* stream) {
  using namespace lrb;
  LRB_REQUIRE(ids && tok_first && tok_offset, "lrb_prepare_sequences: null pointer");
  LRB_REQUIRE(B > 0 && L > 0, "lrb_prepare_sequences: bad shape");
  LRB_REQUIRE(id_bytes == 4 || id_bytes == 8, "lrb_prepare_sequences: id_bytes must be 4 (int32) or 8 (int64)");
  if (L > LRB_MAX_LEN)
    return set_error(LRB_ERR_UNSUPPORTED, "sequence length %d exceeds LRB_MAX_LEN=%d", L, LRB_MAX_LEN);
  LRB_REQUIRE((excl_sorted == nullptr) == (excl_bloom == nullptr), "lrb_prepare_sequences: exclusion list and bloom filter must come together");
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  cudaStream_t st = as_stream(stream);
  prepare_sequences_kernel<<<B, SEQ_THREADS, 0, st>>>(ids, id_bytes, B, L, all_positions, tok_first, tok_offset, excl_sorted,
                                                       excl_bloom, lrb_excl_stride(L));
  LRB_CUDA_TRY(cudaGetLastError());
  scan_counts_kernel<<<1, 1024, 0, st>>>(tok_offset, B);
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}

}  // extern "C"
