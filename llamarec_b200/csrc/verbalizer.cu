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
// HBM/latency-bound skinny GEMM.  UB = 4 users per CTA: their hidden rows are staged once in shared
// memory (32 KB at H = 4096); each of the 8 warps owns a subset of the label words and streams those
// lm_head rows straight from L2 with 16-byte loads (a whole row's loads are issued before the first
// FMA so ~H/256 requests per lane are in flight), dotting them with the 4 staged rows at once.  The
// per-user post-processing (project / softmax / log / aggregate) then runs in one warp per user.
#include "api_util.h"
#include "common.cuh"

namespace lrb {
namespace verb {

constexpr int WARPS = 8;
constexpr int UB = 4;           // users per CTA
constexpr int MAX_WORDS = 32;   // C * W upper bound
constexpr int MAX_VEC = 32;     // 16-byte vectors per lane per row => H <= 8192

struct Params {
  const __nv_bfloat16* hidden;   // [B][H]
  const __nv_bfloat16* lm_head;  // [V][H]
  int B, H;
  long long V;
  const int* word_ids;           // [C][W]
  const unsigned char* word_mask;// [C][W]
  int C, W, mode, round_bf16;
  const float* calib_logits;     // optional [V] fp32: ManualVerbalizer._calibrate_logits (trainer/verb.py:202-208)
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

// Verbalizer post-processing for one user, executed by a whole warp (lanes = label words):
//   project   : logit - 10000 * (1 - word_mask)                          trainer/verb.py:543
//   normalize : softmax over ALL label words, log(p + 1e-15)   (mode 1)  trainer/verb.py:570-582
//   calibrate : (mode 1, optional) p /= softmax(project(calibration logits)) + 1e-15, renormalised over all
//               label words                                             trainer/verb.py:616-643
//   aggregate : masked mean over the W words of each class               trainer/verb.py:611-614
// calib_logit = this lane's label-word logit of the calibration vector (same multi-token handler), unused when
// has_calib is false.
LRB_DEVINL void verbalizer_tail(int mode, int C, int W, const unsigned char* word_mask, float logit, int lane,
                                float* out_row, bool has_calib = false, float calib_logit = 0.f) {
  const int n_words = C * W;
  float x = -INFINITY;
  float m = 0.f;
  if (lane < n_words) {
    m = word_mask[lane] ? 1.f : 0.f;
    x = logit - 10000.0f * (1.0f - m);
  }
  if (mode == 1) {
    float mx = x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = lane < n_words ? expf(x - mx) : 0.f;
    const float den = warp_sum(e);
    float pr = e / den;
    if (has_calib) {
      const float cx = lane < n_words ? calib_logit - 10000.0f * (1.0f - m) : -INFINITY;
      float cmx = cx;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cmx = fmaxf(cmx, __shfl_xor_sync(0xffffffffu, cmx, o));
      const float ce = lane < n_words ? expf(cx - cmx) : 0.f;
      const float cden = warp_sum(ce);
      pr = pr / (ce / cden + 1e-15f);
      const float norm = warp_sum(lane < n_words ? pr : 0.f);
      pr = pr / norm;
    }
    x = logf(pr + 1e-15f);
  }
  const float num = lane < n_words ? x * m : 0.f;
  float tot = 0.f, totm = 0.f;
  for (int j = 0; j < W; ++j) {
    const int src = (lane / W) * W + j;
    tot += __shfl_sync(0xffffffffu, num, src & 31);
    totm += __shfl_sync(0xffffffffu, m, src & 31);
  }
  if (lane < n_words && (lane % W) == 0) out_row[lane / W] = tot / totm;
}

template <int NVEC>   // NVEC = H / 256: 16-byte vectors per lane per row
__global__ void __launch_bounds__(WARPS * 32) verbalizer_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* sH = reinterpret_cast<uint4*>(smem_raw);        // [UB][H/8] 16-byte vectors
  __shared__ float s_logit[UB][MAX_WORDS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * UB;
  const int n_words = p.C * p.W;
  const int hv = p.H / 8;                                 // vectors per row

  for (int e = threadIdx.x; e < UB * hv; e += WARPS * 32) {
    const int u = e / hv, v = e - u * hv;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (b0 + u < p.B) val = __ldg(reinterpret_cast<const uint4*>(p.hidden + static_cast<size_t>(b0 + u) * p.H) + v);
    sH[e] = val;
  }
  __syncthreads();

  for (int w = warp; w < n_words; w += WARPS) {
    long long tok = p.word_ids[w];
    if (tok < 0 || tok >= p.V) tok = 0;
    const uint4* row = reinterpret_cast<const uint4*>(p.lm_head + static_cast<size_t>(tok) * p.H);
    uint4 wv[NVEC];
#pragma unroll
    for (int j = 0; j < NVEC; ++j) wv[j] = __ldg(row + j * 32 + lane);   // all loads in flight
    float acc[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) acc[u] = 0.f;
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      float wf[8];
      bf16x8_to_f32(wv[j], wf);
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        float hf[8];
        bf16x8_to_f32(sH[u * hv + j * 32 + lane], hf);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[u] = fmaf(hf[i], wf[i], acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      float s = warp_sum(acc[u]);
      if (p.round_bf16) s = __bfloat162float(__float2bfloat16_rn(s));
      if (lane == 0) s_logit[u][w] = s;
    }
  }
  __syncthreads();
  if (warp >= UB) return;
  const int b = b0 + warp;
  if (b >= p.B) return;

  float cal = 0.f;
  if (p.calib_logits != nullptr && lane < n_words) {
    long long tok = p.word_ids[lane];
    if (tok < 0 || tok >= p.V) tok = 0;
    cal = __ldg(p.calib_logits + tok);
  }
  verbalizer_tail(p.mode, p.C, p.W, p.word_mask, lane < n_words ? s_logit[warp][lane] : 0.f, lane,
                  p.out + static_cast<size_t>(b) * p.C, p.calib_logits != nullptr, cal);
}

// ---------------------------------------------------------------------------------------------
// process_logits on precomputed logits [B][V] (the reference's own entry point, trainer/verb.py:546-586):
// one warp per user, lanes = label words; the word's sub-token logits are gathered and reduced by the
// multi_token_handler (first / max / mean, trainer/verb.py:280-305), then the same tail runs.
// ---------------------------------------------------------------------------------------------
struct LogitParams {
  const float* logits;             // [B][V]
  long long ld;                    // row pitch of logits (elements)
  int B;
  long long V;
  const int* tok_ids;              // [C][W][T]
  const unsigned char* tok_mask;   // [C][W][T]
  const unsigned char* word_mask;  // [C][W]
  int C, W, T, mode, handler;      // handler 0 = first, 1 = max, 2 = mean
  const float* calib_logits;       // optional [V] fp32 calibration logits
  float* out;                      // [B][C]
};

__global__ void __launch_bounds__(WARPS * 32) verbalizer_logits_kernel(const LogitParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= p.B) return;
  const int n_words = p.C * p.W;
  // this lane's label-word logit of one logits row: sub-tokens gathered and reduced by the handler
  auto word_logit = [&](const float* row) {
    const int* ids = p.tok_ids + lane * p.T;
    const unsigned char* tm = p.tok_mask + lane * p.T;
    auto at = [&](int t) {
      long long tok = ids[t];
      if (tok < 0 || tok >= p.V) tok = 0;
      return __ldg(row + tok);
    };
    if (p.handler == 0) return at(0);
    if (p.handler == 1) {
      float mx = -INFINITY;
      for (int t = 0; t < p.T; ++t) mx = fmaxf(mx, at(t) - 1000.0f * (1.0f - (tm[t] ? 1.f : 0.f)));
      return mx;
    }
    float sum = 0.f, cnt = 0.f;
    for (int t = 0; t < p.T; ++t) {
      const float m = tm[t] ? 1.f : 0.f;
      sum += at(t) * m;
      cnt += m;
    }
    return sum / (cnt + 1e-15f);
  };
  float x = 0.f, cal = 0.f;
  if (lane < n_words) {
    x = word_logit(p.logits + static_cast<size_t>(b) * p.ld);
    if (p.calib_logits != nullptr) cal = word_logit(p.calib_logits);
  }
  verbalizer_tail(p.mode, p.C, p.W, p.word_mask, x, lane, p.out + static_cast<size_t>(b) * p.C,
                  p.calib_logits != nullptr, cal);
}

}  // namespace verb
}  // namespace lrb

extern "C" int lrb_verbalizer_score(const void* hidden_bf16, const void* lm_head_bf16, int B, int H, int64_t V,
                                    const int32_t* word_ids, const uint8_t* word_mask, int C, int W, int mode,
                                    int round_bf16, const float* calib_logits, float* out, void* stream) {
  using namespace lrb;
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  LRB_REQUIRE(hidden_bf16 && lm_head_bf16 && word_ids && word_mask && out, "lrb_verbalizer_score: null pointer");
  LRB_REQUIRE(B > 0 && V > 0 && C > 0 && W > 0, "lrb_verbalizer_score: bad shape");
  LRB_REQUIRE(mode == 0 || mode == 1, "lrb_verbalizer_score: mode must be 0 (raw) or 1 (log-softmax)");
  if (H % 256 != 0 || H / 256 > verb::MAX_VEC)
    return set_error(LRB_ERR_UNSUPPORTED, "hidden size %d must be a multiple of 256 and <= %d", H, 256 * verb::MAX_VEC);
  if (C * W > verb::MAX_WORDS)
    return set_error(LRB_ERR_UNSUPPORTED, "at most %d label words in total (got %d x %d)", verb::MAX_WORDS, C, W);
  verb::Params p;
  p.hidden = static_cast<const __nv_bfloat16*>(hidden_bf16);
  p.lm_head = static_cast<const __nv_bfloat16*>(lm_head_bf16);
  p.B = B; p.H = H; p.V = V; p.word_ids = word_ids; p.word_mask = word_mask;
  p.C = C; p.W = W; p.mode = mode; p.round_bf16 = round_bf16; p.calib_logits = calib_logits; p.out = out;
  const size_t smem = static_cast<size_t>(verb::UB) * H * 2;
  const int grid = (B + verb::UB - 1) / verb::UB;
  const int nvec = H / 256;
  cudaStream_t st = as_stream(stream);
#define LRB_VERB_LAUNCH(NV)                                                                              \
  do {                                                                                                   \
    LRB_CUDA_TRY(cudaFuncSetAttribute(verb::verbalizer_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      static_cast<int>(smem)));                                          \
    verb::verbalizer_kernel<NV><<<grid, verb::WARPS * 32, smem, st>>>(p);                                \
  } while (0)
  if (nvec <= 2) { if (nvec == 1) LRB_VERB_LAUNCH(1); else LRB_VERB_LAUNCH(2); }
  else if (nvec <= 4) { if (nvec == 3) LRB_VERB_LAUNCH(3); else LRB_VERB_LAUNCH(4); }
  else if (nvec == 6) LRB_VERB_LAUNCH(6);
  else if (nvec == 8) LRB_VERB_LAUNCH(8);
  else if (nvec == 12) LRB_VERB_LAUNCH(12);
  else if (nvec == 14) LRB_VERB_LAUNCH(14);
  else if (nvec == 16) LRB_VERB_LAUNCH(16);
  else if (nvec == 20) LRB_VERB_LAUNCH(20);
  else if (nvec == 32) LRB_VERB_LAUNCH(32);
  else return set_error(LRB_ERR_UNSUPPORTED, "hidden size %d is not one of the instantiated widths", H);
#undef LRB_VERB_LAUNCH
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}

extern "C" int lrb_verbalizer_from_logits(const float* logits, int64_t ld, int B, int64_t V, const int32_t* tok_ids,
                                          const uint8_t* tok_mask, const uint8_t* word_mask, int C, int W, int T,
                                          int handler, int mode, const float* calib_logits, float* out, void* stream) {
  using namespace lrb;
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  LRB_REQUIRE(logits && tok_ids && tok_mask && word_mask && out, "lrb_verbalizer_from_logits: null pointer");
  LRB_REQUIRE(B > 0 && V > 0 && C > 0 && W > 0 && T > 0 && ld >= V, "lrb_verbalizer_from_logits: bad shape");
  LRB_REQUIRE(mode == 0 || mode == 1, "lrb_verbalizer_from_logits: mode must be 0 (raw) or 1 (log-softmax)");
  LRB_REQUIRE(handler >= 0 && handler <= 2, "lrb_verbalizer_from_logits: handler is 0 first, 1 max, 2 mean");
  if (C * W > verb::MAX_WORDS)
    return set_error(LRB_ERR_UNSUPPORTED, "at most %d label words in total (got %d x %d)", verb::MAX_WORDS, C, W);
  verb::LogitParams p;
  p.logits = logits; p.ld = ld; p.B = B; p.V = V; p.tok_ids = tok_ids; p.tok_mask = tok_mask;
  p.word_mask = word_mask; p.C = C; p.W = W; p.T = T; p.mode = mode; p.handler = handler;
  p.calib_logits = calib_logits; p.out = out;
  const int grid = (B + verb::WARPS - 1) / verb::WARPS;
  verb::verbalizer_logits_kernel<<<grid, verb::WARPS * 32, 0, as_stream(stream)>>>(p);
  LRB_CUDA_TRY(cudaGetLastError());
  return LRB_OK;
}
