// Stage-2 verbalizer tail (north_star subsystem 4): only the label-word rows of the lm_head are
// multiplied with the last-position hidden state, then the verbalizer post-processing runs in the
// same kernel.
//
// Replaces   logits = lm_head(hidden_states).float()[:, -1]      model/llm.py:113-114,131
//            ManualVerbalizer.process_logits(logits)            trainer/verb.py:546-586
//              project   : logits[:, label_words_ids], first sub-token, -10000*(1-mask)   :539-544
//              normalize : softmax over all label words (post_log_softmax only)           :588-600
//              log(p + 1e-15), aggregate = masked mean over the words of a class          :582,611-614
//
// HBM-bound skinny GEMM: 8 users per CTA (one warp each); the C*W label rows are staged through
// shared memory in K-chunks of 512 so each row is read once per CTA, not once per user.
#include "api_util.h"
#include "common.cuh"

namespace lrb {
namespace verb {

constexpr int WARPS = 8;
constexpr int KC = 512;        // K chunk (bf16 elements)
constexpr int MAX_WORDS = 32;  // C * W upper bound

struct Params {
  const __nv_bfloat16* hidden;   // [B][H]
  const __nv_bfloat16* lm_head;  // [V][H]
  int B, H;
  long long V;
  const int* word_ids;           // [C][W]
  const unsigned char* word_mask;// [C][W]
  int C, W, mode, round_bf16;
  float* out;                    // [B][C]
};

LRB_DEVINL void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(WARPS * 32) verbalizer_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* sW = reinterpret_cast<uint4*>(smem_raw);   // [n_words][KC/8] 16-byte vectors
  __shared__ float s_logit[WARPS][MAX_WORDS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  const int n_words = p.C * p.W;
  const bool live = b < p.B;

  float acc[MAX_WORDS];
#pragma unroll
  for (int w = 0; w < MAX_WORDS; ++w) acc[w] = 0.f;

  for (int k0 = 0; k0 < p.H; k0 += KC) {
    __syncthreads();
    // stage the label rows' chunk
    for (int e = threadIdx.x; e < n_words * (KC / 8); e += WARPS * 32) {
      const int w = e / (KC / 8);
      const int v = e - w * (KC / 8);
      long long tok = p.word_ids[w];
      if (tok < 0 || tok >= p.V) tok = 0;
      sW[e] = __ldg(reinterpret_cast<const uint4*>(p.lm_head + static_cast<size_t>(tok) * p.H + k0) + v);
    }
    // this user's chunk of the hidden state: two 16-byte vectors per lane
    float h0[8], h1[8];
    {
      uint4 a = make_uint4(0u, 0u, 0u, 0u), c = a;
      if (live) {
        const uint4* hp = reinterpret_cast<const uint4*>(p.hidden + static_cast<size_t>(b) * p.H + k0);
        a = __ldg(hp + lane);
        c = __ldg(hp + 32 + lane);
      }
      bf16x8_to_f32(a, h0);
      bf16x8_to_f32(c, h1);
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < MAX_WORDS; ++w) {
      if (w < n_words) {
        float f0[8], f1[8];
        bf16x8_to_f32(sW[w * (KC / 8) + lane], f0);
        bf16x8_to_f32(sW[w * (KC / 8) + 32 + lane], f1);
        float s = acc[w];
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(h0[i], f0[i], s);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(h1[i], f1[i], s);
        acc[w] = s;
      }
    }
  }
#pragma unroll
  for (int w = 0; w < MAX_WORDS; ++w) {
    if (w < n_words) {
      float s = warp_sum(acc[w]);
      if (p.round_bf16) s = __bfloat162float(__float2bfloat16_rn(s));
      if (lane == 0) s_logit[warp][w] = s;
    }
  }
  __syncwarp();
  if (!live) return;

  // ---- verbalizer post-processing, lanes = label words ----
  float x = -INFINITY;
  float m = 0.f;
  if (lane < n_words) {
    m = p.word_mask[lane] ? 1.f : 0.f;
    x = s_logit[warp][lane] - 10000.0f * (1.0f - m);      // trainer/verb.py:543
  }
  if (p.mode == 1) {
    // softmax over ALL label words of all classes, then log(p + 1e-15)   (trainer/verb.py:570-582)
    float mx = x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = lane < n_words ? expf(x - mx) : 0.f;
    const float den = warp_sum(e);
    x = logf(e / den + 1e-15f);
  }
  // aggregate: masked mean over the W words of each class                   (trainer/verb.py:611-614)
  float num = lane < n_words ? x * m : 0.f;
  float cntm = m;
  // words of class c sit in lanes c*W .. c*W+W-1: segmented sum by walking the W neighbours
  float tot = 0.f, totm = 0.f;
  for (int j = 0; j < p.W; ++j) {
    const int src = (lane / p.W) * p.W + j;
    tot += __shfl_sync(0xffffffffu, num, src & 31);
    totm += __shfl_sync(0xffffffffu, cntm, src & 31);
  }
  if (lane < n_words && (lane % p.W) == 0) p.out[static_cast<size_t>(b) * p.C + lane / p.W] = tot / totm;
}

}  // namespace verb
}  // namespace lrb

extern "C" int lrb_verbalizer_score(const void* hidden_bf16, const void* lm_head_bf16, int B, int H, int64_t V,
                                    const int32_t* word_ids, const uint8_t* word_mask, int C, int W, int mode,
                                    int round_bf16, float* out, void* stream) {
  using namespace lrb;
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  LRB_REQUIRE(hidden_bf16 && lm_head_bf16 && word_ids && word_mask && out, "lrb_verbalizer_score: null pointer");
  LRB_REQUIRE(B > 0 && V > 0 && C > 0 && W > 0, "lrb_verbalizer_score: bad shape");
  LRB_REQUIRE(mode == 0 || mode == 1, "lrb_verbalizer_score: mode must be 0 (raw) or 1 (log-softmax)");
  if (H % verb::KC != 0)
    return set_error(LRB_ERR_UNSUPPORTED, "hidden size %d must be a multiple of %d", H, verb::KC);
  if (C * W > verb::MAX_WORDS)
    return set_error(LRB_ERR_UNSUPPORTED, "at most %d label words in total (got %d x %d)", verb::MAX_WORDS, C, W);
  verb::Params p;
  p.hidden = static_cast<const __nv_bfloat16*>(hidden_bf16);
  p.lm_head = static_cast<const __nv_bfloat16*>(lm_head_bf16);
  p.B = B; p.H = H; p.V = V; p.word_ids = word_ids; p.word_mask = word_mask;
  p.C = C; p.W = W; p.mode = mode; p.round_bf16 = round_bf16; p.out = out;
  const size_t smem = static_cast<size_t>(C) * W * verb::KC * 2;
  const int grid = (B + verb::WARPS - 1) / verb::WARPS;
  verb::verbalizer_kernel<<<grid, verb::WARPS * 32, smem, as_stream(stream)>>>(p);
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}
