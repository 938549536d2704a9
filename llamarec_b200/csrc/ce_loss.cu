// Train-step forward loss without the [B*L, N+1] logits tensor (forward half).
//
// Replaces, in value,
//     logits = model(seqs).view(-1, N+1)                 trainer/lru.py:24-25   (4.5 GB at Games shape, batch 2048)
//     loss   = CrossEntropyLoss(ignore_index=0)(logits, labels.view(-1))   trainer/lru.py:20,26-27
// by an online log-sum-exp over the catalogue fused with the scoring contraction: per row only
// (running max, running sum of exp, the label's logit) leave the SM.
//
// fp32 FFMA path (training catalogues are ~10^4 items; the loss must agree with the fp32 reference to ~1e-6):
// thread = hidden row (64 values in registers), CTA = 128 rows, item chunks of 64 rows staged in shared
// memory; gridDim.y splits the item range, a second tiny kernel combines the splits and reduces the mean.
#include "api_util.h"
#include "common.cuh"

#include <climits>
#include <cmath>

namespace lrb {
namespace ce {

constexpr int ROWS = 128;
constexpr int IT = 64;
constexpr int D = 64;

struct Params {
  const float* x;          // [M][64] hidden states at every position
  const float* table;      // [rows][64]
  const float* bias_pad;   // [ceil(rows/256)*256]
  const long long* labels; // [M]
  int M, rows, splits;
  float* part_max;         // [M][splits]
  float* part_sum;         // [M][splits]
  float* label_logit;      // [M]
};

__global__ void __launch_bounds__(ROWS) ce_partial_kernel(const Params p) {
  __shared__ __align__(16) float sE[IT * D];
  __shared__ float sBias[IT];
  const int tid = threadIdx.x;
  const int m = blockIdx.x * ROWS + tid;
  const bool live = m < p.M;
  float u[D];
#pragma unroll
  for (int k = 0; k < D; k += 4) {
    float4 v = live ? *reinterpret_cast<const float4*>(p.x + static_cast<size_t>(m) * D + k)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    u[k] = v.x; u[k + 1] = v.y; u[k + 2] = v.z; u[k + 3] = v.w;
  }
  const long long label = live ? p.labels[m] : -1;
  const int chunks = (p.rows + IT - 1) / IT;
  const int c0 = static_cast<int>((static_cast<long long>(blockIdx.y) * chunks) / gridDim.y);
  const int c1 = static_cast<int>((static_cast<long long>(blockIdx.y + 1) * chunks) / gridDim.y);
  float run_max = -INFINITY, run_sum = 0.f;
  for (int c = c0; c < c1; ++c) {
    const int i0 = c * IT;
    __syncthreads();
    for (int e = tid; e < IT * D / 4; e += ROWS) {
      const int it = e / (D / 4);
      const int k4 = e - it * (D / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + it < p.rows) v = *reinterpret_cast<const float4*>(p.table + static_cast<size_t>(i0 + it) * D + k4 * 4);
      reinterpret_cast<float4*>(sE)[e] = v;
    }
    if (tid < IT) sBias[tid] = p.bias_pad[i0 + tid];   // -inf beyond rows: exp() contributes 0
    __syncthreads();
    // 16 scores at a time (registers), then one rescale of the running sum per 16 items (not per item)
    const int lab_off = (label >= i0 && label < i0 + IT && label < p.rows) ? static_cast<int>(label - i0) : -1;
    for (int g = 0; g < IT / 16; ++g) {
      float s[16];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4* e4 = reinterpret_cast<const float4*>(sE + (g * 16 + j) * D);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // same summation order as score_f32_kernel
#pragma unroll
        for (int k4 = 0; k4 < D / 4; ++k4) {
          const float4 ev = e4[k4];
          a0 = fmaf(u[4 * k4 + 0], ev.x, a0);
          a1 = fmaf(u[4 * k4 + 1], ev.y, a1);
          a2 = fmaf(u[4 * k4 + 2], ev.z, a2);
          a3 = fmaf(u[4 * k4 + 3], ev.w, a3);
        }
        s[j] = ((a0 + a1) + (a2 + a3)) + sBias[g * 16 + j];
        cmax = fmaxf(cmax, s[j]);
        if (g * 16 + j == lab_off && live) p.label_logit[m] = s[j];
      }
      if (cmax > -INFINITY) {
        const float nmax = fmaxf(run_max, cmax);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += expf(s[j] - nmax);
        run_sum = run_sum * expf(run_max - nmax) + acc;   // exp(-inf) = 0 on the first group
        run_max = nmax;
      }
    }
  }
  if (live) {
    p.part_max[static_cast<size_t>(m) * p.splits + blockIdx.y] = run_max;
    p.part_sum[static_cast<size_t>(m) * p.splits + blockIdx.y] = run_sum;
  }
}

struct FinParams {
  const float* part_max;
  const float* part_sum;
  const float* label_logit;
  const long long* labels;
  long long ignore_index;
  int M, splits, rows;
  float* row_loss;     // [M] (0 for ignored rows), may be null
  float* loss_sum;     // [2]: sum of row losses, number of counted rows (accumulated)
};

__global__ void __launch_bounds__(256) ce_finalize_kernel(const FinParams p) {
  __shared__ float s_sum[8], s_cnt[8];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f, cnt = 0.f;
  if (m < p.M) {
    const long long label = p.labels[m];
    if (label != p.ignore_index && label >= 0 && label < p.rows) {
      float mx = -INFINITY;
      for (int j = 0; j < p.splits; ++j) mx = fmaxf(mx, p.part_max[static_cast<size_t>(m) * p.splits + j]);
      float sum = 0.f;
      for (int j = 0; j < p.splits; ++j)
        sum += p.part_sum[static_cast<size_t>(m) * p.splits + j] *
               expf(p.part_max[static_cast<size_t>(m) * p.splits + j] - mx);
      loss = (mx + logf(sum)) - p.label_logit[m];
      cnt = 1.f;
    }
    if (p.row_loss) p.row_loss[m] = loss;
  }
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = loss; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += s_sum[w]; b += s_cnt[w]; }
    if (b != 0.f) {
      atomicAdd(p.loss_sum, a);
      atomicAdd(p.loss_sum + 1, b);
    }
  }
}

inline int splits_for(long long M, long long rows, int sms) {
  const long long m_tiles = (M + ROWS - 1) / ROWS;
  const long long chunks = (rows + IT - 1) / IT;
  long long want = (4LL * sms + m_tiles - 1) / m_tiles;
  const long long max_splits = chunks / 8 > 0 ? chunks / 8 : 1;   // >= 512 items per split
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return static_cast<int>(want);
}

}  // namespace ce
}  // namespace lrb

extern "C" {

size_t lrb_ce_workspace_bytes(int64_t M, int64_t rows) {
  int sms = lrb::device_sm_count();
  if (sms <= 0) sms = 148;
  if (M < 1 || rows < 1) return 0;
  const size_t splits = static_cast<size_t>(lrb::ce::splits_for(M, rows, sms));
  return static_cast<size_t>(M) * (2 * splits + 1) * sizeof(float) + 256;
}

int lrb_ce_loss_fwd(const float* hidden, const float* table_f32, const float* bias_pad, int64_t M, int64_t rows,
                    const int64_t* labels, int64_t ignore_index, float* row_loss, float* loss_sum,
                    void* workspace, size_t workspace_bytes, void* stream) {
  using namespace lrb;
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  LRB_REQUIRE(hidden && table_f32 && bias_pad && labels && loss_sum && workspace, "lrb_ce_loss_fwd: null pointer");
  LRB_REQUIRE(M > 0 && M < INT_MAX && rows > 0 && rows < INT_MAX, "lrb_ce_loss_fwd: bad shape");
  if (workspace_bytes < lrb_ce_workspace_bytes(M, rows))
    return set_error(LRB_ERR_WORKSPACE, "lrb_ce_loss_fwd: workspace too small");
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int splits = ce::splits_for(M, rows, sms);
  float* ws = static_cast<float*>(workspace);
  ce::Params p;
  p.x = hidden; p.table = table_f32; p.bias_pad = bias_pad;
  p.labels = reinterpret_cast<const long long*>(labels);
  p.M = static_cast<int>(M); p.rows = static_cast<int>(rows); p.splits = splits;
  p.part_max = ws;
  p.part_sum = ws + static_cast<size_t>(M) * splits;
  p.label_logit = ws + 2 * static_cast<size_t>(M) * splits;
  cudaStream_t st = as_stream(stream);
  dim3 grid(static_cast<unsigned>((M + ce::ROWS - 1) / ce::ROWS), static_cast<unsigned>(splits));
  ce::ce_partial_kernel<<<grid, ce::ROWS, 0, st>>>(p);
  LRB_CUDA_TRY(cudaGetLastError());
  ce::FinParams f;
  f.part_max = p.part_max; f.part_sum = p.part_sum; f.label_logit = p.label_logit;
  f.labels = p.labels; f.ignore_index = ignore_index; f.M = p.M; f.splits = splits; f.rows = p.rows;
  f.row_loss = row_loss; f.loss_sum = loss_sum;
  ce::ce_finalize_kernel<<<(p.M + 255) / 256, 256, 0, st>>>(f);
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}

}  // extern "C"
