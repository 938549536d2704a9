// Developer harness (not part of the product path): exercises lrb_score_dense / lrb_score_topk on a
// real B200 against a brute-force host loop, and times the scoring kernel at a configurable size.
//   build:  make -C llamarec_b200/csrc tc_check       run:  tools/tc_check [B rows K]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../include/llamarec_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)
#define LK(x)                                                        \
  do {                                                               \
    int r_ = (x);                                                    \
    if (r_ != 0) {                                                   \
      printf("lrb error %d: %s (%s)\n", r_, lrb_last_error(), #x);   \
      exit(3);                                                       \
    }                                                                \
  } while (0)

extern "C" void lrb_debug_set_score_mode(int m);
extern "C" void lrb_debug_set_scout(int t);
extern "C" void lrb_debug_set_probe_out(long long* p);
extern "C" void lrb_debug_set_pair_drain(int v);
extern "C" void lrb_debug_set_overlap(int v);
extern "C" void lrb_debug_set_cap_div(int v);
extern "C" void lrb_debug_set_restart_tiles(double v);

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Problem {
  int B, rows, K, L;
  std::vector<float> u, e, bias;           // bf16-rounded values kept in fp32
  std::vector<__nv_bfloat16> u16, e16;
  std::vector<float> bias_pad;
  std::vector<int> excl;                   // [B][stride]
  std::vector<uint32_t> bloom;             // [B][4]
  int stride;
};

static Problem make_problem(int B, int rows, int K, int L, unsigned seed, bool with_bias) {
  Problem p;
  p.B = B; p.rows = rows; p.K = K; p.L = L;
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  p.u.resize((size_t)B * 64); p.u16.resize(p.u.size());
  p.e.resize((size_t)rows * 64); p.e16.resize(p.e.size());
  for (size_t i = 0; i < p.u.size(); ++i) { p.u[i] = bf16_round(nd(rng)); p.u16[i] = __float2bfloat16(p.u[i]); }
  for (size_t i = 0; i < p.e.size(); ++i) { p.e[i] = bf16_round(0.05f * nd(rng)); p.e16[i] = __float2bfloat16(p.e[i]); }
  p.bias.resize(rows);
  for (int i = 0; i < rows; ++i) p.bias[i] = with_bias ? 0.05f * nd(rng) : 0.f;
  int npad = (rows + 255) / 256 * 256;
  p.bias_pad.assign(npad, -INFINITY);
  std::copy(p.bias.begin(), p.bias.end(), p.bias_pad.begin());
  p.stride = lrb_excl_stride(L);
  p.excl.assign((size_t)B * p.stride, 0x7fffffff);
  p.bloom.assign((size_t)B * 4, 0u);
  std::uniform_int_distribution<int> idd(1, rows - 1);
  for (int b = 0; b < B; ++b) {
    std::vector<int> h;
    h.push_back(0);
    int n = 1 + (int)(rng() % L);
    for (int i = 0; i < n; ++i) h.push_back(idd(rng));
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    for (size_t i = 0; i < h.size(); ++i) {
      p.excl[(size_t)b * p.stride + i] = h[i];
      p.bloom[(size_t)b * 4 + ((h[i] >> 5) & 3)] |= 1u << (h[i] & 31);
    }
  }
  return p;
}

static float host_score(const Problem& p, int b, int n) {
  float acc = 0.f;
  for (int k = 0; k < 64; ++k) acc += p.u[(size_t)b * 64 + k] * p.e[(size_t)n * 64 + k];
  return acc + p.bias[n];
}

template <class T>
static T* dev_copy(const std::vector<T>& v) {
  T* d;
  CK(cudaMalloc(&d, v.size() * sizeof(T)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct DevTable {
  void* e16 = nullptr;
  float* bias_pad = nullptr;
  void* bias_blk = nullptr;
  float* e32 = nullptr;
};
static DevTable prepare_dev_table(const std::vector<float>& e32, const std::vector<float>& bias, int rows) {
  DevTable t;
  t.e32 = dev_copy(e32);
  float* dbias = dev_copy(bias);
  CK(cudaMalloc(&t.e16, (size_t)rows * 64 * 2));
  CK(cudaMalloc(&t.bias_pad, (size_t)((rows + 255) / 256 * 256) * 4));
  CK(cudaMalloc(&t.bias_blk, lrb_bias_blk_bytes(rows)));
  LK(lrb_prepare_table(t.e32, dbias, 0, rows, t.e16, t.bias_pad, t.bias_blk, nullptr));
  CK(cudaDeviceSynchronize());
  cudaFree(dbias);
  return t;
}
static void free_dev_table(DevTable& t) { cudaFree(t.e16); cudaFree(t.bias_pad); cudaFree(t.bias_blk); cudaFree(t.e32); }

static int check_dense(int B, int rows) {
  Problem p = make_problem(B, rows, 1, 8, 123, true);
  auto* du = dev_copy(p.u16);
  DevTable T = prepare_dev_table(p.e, p.bias, rows);
  void* de = T.e16;
  float* db = T.bias_pad;
  float* dout;
  long long ld = rows;
  CK(cudaMalloc(&dout, (size_t)B * ld * 4));
  CK(cudaMemset(dout, 0xff, (size_t)B * ld * 4));
  LK(lrb_score_dense(du, de, db, T.bias_blk, B, rows, 0, dout, ld, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)B * ld);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxd = 0;
  int bad = 0;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < rows; ++n) {
      float ref = host_score(p, b, n);
      float got = out[(size_t)b * ld + n];
      double d = std::fabs((double)ref - got);
      if (!(d <= 1e-3 * std::max(1.0, (double)std::fabs(ref)))) {
        if (bad < 8) printf("  dense mismatch b=%d n=%d ref=%f got=%f\n", b, n, ref, got);
        ++bad;
      }
      if (d > maxd) maxd = d;
    }
  printf("[dense bf16] B=%d rows=%d max|diff|=%.3e mismatches=%d\n", B, rows, maxd, bad);
  // fp32 path
  auto* du32 = dev_copy(p.u);
  float* de32 = T.e32;
  CK(cudaMemset(dout, 0xff, (size_t)B * ld * 4));
  LK(lrb_score_dense(du32, de32, db, nullptr, B, rows, 1, dout, ld, nullptr));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxd2 = 0;
  int bad2 = 0;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < rows; ++n) {
      float ref = host_score(p, b, n);
      float got = out[(size_t)b * ld + n];
      double d = std::fabs((double)ref - got);
      if (!(d <= 1e-5 * std::max(1.0, (double)std::fabs(ref)))) ++bad2;
      if (d > maxd2) maxd2 = d;
    }
  printf("[dense fp32] B=%d rows=%d max|diff|=%.3e mismatches=%d\n", B, rows, maxd2, bad2);
  cudaFree(du); free_dev_table(T); cudaFree(dout); cudaFree(du32);
  return bad + bad2;
}

struct Cand { float s; int id; };
static bool cand_better(const Cand& a, const Cand& b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }

static int check_topk(int B, int rows, int K, int precision, int row_offset, bool with_bias = true) {
  Problem p = make_problem(B, rows, K, 50, 77, with_bias);
  // shift exclusion ids into the global frame used by the kernel
  Problem q = p;
  for (auto& x : q.excl) if (x != 0x7fffffff) x += row_offset;
  for (int b = 0; b < B; ++b) {
    for (int w = 0; w < 4; ++w) q.bloom[(size_t)b * 4 + w] = 0;
    for (int i = 0; i < q.stride; ++i) {
      int id = q.excl[(size_t)b * q.stride + i];
      if (id != 0x7fffffff) q.bloom[(size_t)b * 4 + ((id >> 5) & 3)] |= 1u << (id & 31);
    }
  }
  void* du = precision == 0 ? (void*)dev_copy(p.u16) : (void*)dev_copy(p.u);
  DevTable T = prepare_dev_table(p.e, p.bias, rows);
  void* de = precision == 0 ? T.e16 : (void*)T.e32;
  float* db = T.bias_pad;
  auto* dex = dev_copy(q.excl);
  auto* dbl = dev_copy(q.bloom);
  int slots = 0;
  LK(lrb_score_topk_slots(B, rows, precision, &slots));
  float* dps; int* dpi; int* dpc; void* scratch;
  CK(cudaMalloc(&dps, (size_t)B * slots * K * 4));
  CK(cudaMalloc(&dpi, (size_t)B * slots * K * 4));
  CK(cudaMalloc(&dpc, (size_t)B * slots * 4));
  CK(cudaMemset(dpc, 0, (size_t)B * slots * 4));
  CK(cudaMalloc(&scratch, lrb_score_scratch_bytes(B)));
  LK(lrb_score_topk(du, de, db, with_bias ? T.bias_blk : nullptr, B, rows, row_offset, dex, dbl, q.stride, K, precision, dps,
                    dpi, dpc, slots, scratch, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<float> ps((size_t)B * slots * K);
  std::vector<int> pi(ps.size()), pc((size_t)B * slots);
  CK(cudaMemcpy(ps.data(), dps, ps.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pi.data(), dpi, pi.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pc.data(), dpc, pc.size() * 4, cudaMemcpyDeviceToHost));
  int bad_users = 0;
  double max_sd = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<Cand> got;
    for (int s = 0; s < slots; ++s) {
      int c = pc[(size_t)b * slots + s];
      for (int i = 0; i < c; ++i) {
        size_t o = ((size_t)b * slots + s) * K + i;
        got.push_back({ps[o], pi[o]});
      }
    }
    std::sort(got.begin(), got.end(), cand_better);
    if ((int)got.size() > K) got.resize(K);
    std::vector<Cand> ref;
    for (int n = 0; n < rows; ++n) {
      const int* ex = &p.excl[(size_t)b * p.stride];
      if (std::binary_search(ex, ex + p.stride, n)) continue;
      ref.push_back({host_score(p, b, n), n + row_offset});
    }
    std::partial_sort(ref.begin(), ref.begin() + std::min<size_t>(K, ref.size()), ref.end(), cand_better);
    ref.resize(std::min<size_t>(K, ref.size()));
    bool ok = got.size() == ref.size();
    for (size_t i = 0; ok && i < ref.size(); ++i) {
      double d = std::fabs((double)got[i].s - ref[i].s);
      if (d > max_sd) max_sd = d;
      if (got[i].id != ref[i].id) {
        // tolerate swaps between scores that agree to 1e-5 (fp32 summation order)
        bool tie = d <= 1e-5 * std::max(1.0, (double)std::fabs(ref[i].s));
        if (!tie) ok = false;
      }
    }
    if (!ok) {
      if (bad_users < 4) {
        printf("  topk mismatch user %d (got %zu, ref %zu):\n", b, got.size(), ref.size());
        for (size_t i = 0; i < std::min<size_t>(6, ref.size()); ++i)
          printf("    ref (%f,%d)  got (%f,%d)\n", ref[i].s, ref[i].id,
                 i < got.size() ? got[i].s : 0.f, i < got.size() ? got[i].id : -9);
      }
      ++bad_users;
    }
  }
  printf("[topk prec=%d] B=%d rows=%d K=%d slots=%d off=%d bad_users=%d max|ds|=%.3e\n", precision, B,
         rows, K, slots, row_offset, bad_users, max_sd);
  cudaFree(du); free_dev_table(T); cudaFree(dex); cudaFree(dbl);
  cudaFree(dps); cudaFree(dpi); cudaFree(dpc); cudaFree(scratch);
  return bad_users;
}

static void time_topk(int B, int rows, int K, bool with_bias, bool with_excl) {
  std::mt19937 rng(5);
  std::vector<__nv_bfloat16> u16((size_t)B * 64);
  std::normal_distribution<float> nd(0.f, 1.f);
  for (auto& x : u16) x = __float2bfloat16(nd(rng));
  std::vector<float> e32((size_t)rows * 64);
  {
    uint32_t s = 12345;   // cheap LCG gaussian-ish fill for the big table
    for (auto& x : e32) {
      float a = 0;
      for (int t = 0; t < 4; ++t) { s = s * 1664525u + 1013904223u; a += (s >> 8) * (1.0f / 16777216.0f); }
      x = 0.02f * (a - 2.0f) * 1.73f;
    }
  }
  std::vector<float> bias(rows);
  for (int i = 0; i < rows; ++i) bias[i] = with_bias ? 0.01f * nd(rng) : 0.f;
  DevTable T = prepare_dev_table(e32, bias, rows);
  cudaFree(T.e32); T.e32 = nullptr;
  int L = 50, stride = lrb_excl_stride(L);
  std::vector<int> excl((size_t)B * stride, 0x7fffffff);
  std::vector<uint32_t> bloom((size_t)B * 4, 0u);
  for (int b = 0; b < B; ++b) {
    std::vector<int> h{0};
    for (int i = 0; i < 9; ++i) h.push_back(1 + rng() % (rows - 1));
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    for (size_t i = 0; i < h.size(); ++i) {
      excl[(size_t)b * stride + i] = h[i];
      bloom[(size_t)b * 4 + ((h[i] >> 5) & 3)] |= 1u << (h[i] & 31);
    }
  }
  auto* du = dev_copy(u16);
  auto* dex = dev_copy(excl); auto* dbl = dev_copy(bloom);
  int slots = 0;
  LK(lrb_score_topk_slots(B, rows, 0, &slots));
  float* dps; int* dpi; int* dpc; void* scratch;
  CK(cudaMalloc(&dps, (size_t)B * slots * K * 4));
  CK(cudaMalloc(&dpi, (size_t)B * slots * K * 4));
  CK(cudaMalloc(&dpc, (size_t)B * slots * 4));
  CK(cudaMalloc(&scratch, lrb_score_scratch_bytes(B)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const void* bblk = with_bias ? T.bias_blk : nullptr;
  for (int it = 0; it < 3; ++it)
    LK(lrb_score_topk(du, T.e16, T.bias_pad, bblk, B, rows, 0, with_excl ? dex : nullptr, with_excl ? dbl : nullptr,
                      stride, K, 0, dps, dpi, dpc, slots, scratch, nullptr));
  CK(cudaDeviceSynchronize());
  const int iters = 10;
  CK(cudaEventRecord(e0));
  for (int it = 0; it < iters; ++it)
    LK(lrb_score_topk(du, T.e16, T.bias_pad, bblk, B, rows, 0, with_excl ? dex : nullptr, with_excl ? dbl : nullptr,
                      stride, K, 0, dps, dpi, dpc, slots, scratch, nullptr));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  {
    // one more launch with the per-CTA timers (probe builds only write them): SM cycles and wall time per CTA
    int sms = 0, cc = 0;
    lrb_device_info(&sms, &cc);
    long long* dprobe;
    CK(cudaMalloc(&dprobe, (sms * 8 + 16 * 24) * sizeof(long long)));
    CK(cudaMemset(dprobe, 0, (sms * 8 + 16 * 24) * sizeof(long long)));
    lrb_debug_set_probe_out(dprobe);
    LK(lrb_score_topk(du, T.e16, T.bias_pad, bblk, B, rows, 0, with_excl ? dex : nullptr, with_excl ? dbl : nullptr,
                      stride, K, 0, dps, dpi, dpc, slots, scratch, nullptr));
    CK(cudaDeviceSynchronize());
    lrb_debug_set_probe_out(nullptr);
    std::vector<long long> st(sms * 8 + 16 * 24);
    CK(cudaMemcpy(st.data(), dprobe, st.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double cyc = 0, ns = 0, ew = 0, me = 0, dc = 0, dn = 0, ap = 0; int n = 0, nl = 0;
    for (int c = 0; c < sms; ++c) if (st[c * 8] > 0) {
      cyc += st[c * 8]; ns += st[c * 8 + 1]; ew += st[c * 8 + 2]; dc += st[c * 8 + 5]; dn += st[c * 8 + 6]; ap += st[c * 8 + 7]; ++n;
      if (st[c * 8 + 3] > 0) { me += st[c * 8 + 3]; ++nl; }
    }
    if (n > 0)
      printf("  per CTA: %.3f Mcycles in %.3f ms => SM clock %.0f MHz | epilogue warp 0: waited %.3f Mcycles for accumulators, "
             "%.1f drains taking %.3f Mcycles, thread 0 appended %.1f records | MMA warp 1 waited %.3f Mcycles for operands / "
             "a free accumulator stage\n",
             cyc / n / 1e6, ns / n / 1e6, cyc / ns * 1e3, ew / n / 1e6, dn / n, dc / n / 1e6, ap / n, nl ? me / nl / 1e6 : 0.0);
    if (getenv("LRB_TIMELINE") && st[sms * 8] > 0) {
      const long long* tl = st.data() + sms * 8;
      const long long base = tl[0];
      printf("  timeline of CTA 0 (cycles since the MMA warp woke up for tile T0): per tile  mma_wake mma_commit | done seen by epilogue warps 0-7 | released by warps 0-7\n");
      for (int i = 0; i < 16; ++i) {
        printf("   t+%2d %6lld %6lld (ops %6lld stage %6lld) |", i, tl[i * 24] - base, tl[i * 24 + 1] - base, tl[i * 24 + 18] - base,
               tl[i * 24 + 19] - base);
        // (a warp stamps only the half-tile it owns: warps 0-3 see the first half-tile "done", warps 4-7 release the second)
        for (int w = 0; w < 8; ++w) { if (tl[i * 24 + 2 + w]) printf(" %6lld", tl[i * 24 + 2 + w] - base); else printf("      -"); }
        printf(" |");
        for (int w = 0; w < 8; ++w) { if (tl[i * 24 + 10 + w]) printf(" %6lld", tl[i * 24 + 10 + w] - base); else printf("      -"); }
        printf("\n");
      }
    }
    cudaFree(dprobe);
  }
  double tflops = 2.0 * B * (double)rows * 64 / (ms * 1e-3) / 1e12;
  printf("[time] B=%d rows=%d K=%d bias=%d excl=%d slots=%d : %.3f ms/iter  %.1f TFLOP/s  %.0f users/s\n", B, rows, K,
         (int)with_bias, (int)with_excl, slots, ms, tflops, B / (ms * 1e-3));
  cudaFree(du); free_dev_table(T); cudaFree(dex); cudaFree(dbl);
  cudaFree(dps); cudaFree(dpi); cudaFree(dpc); cudaFree(scratch);
}

int main(int argc, char** argv) {
  int sms = 0, cc = 0;
  LK(lrb_device_info(&sms, &cc));
  printf("device: %d SMs, sm_%d\n", sms, cc);
  int fails = 0;
  if (argc >= 2 && !strcmp(argv[1], "time")) {
    int B = argc > 2 ? atoi(argv[2]) : 4096;
    int rows = argc > 3 ? atoi(argv[3]) : 2500000;
    int K = argc > 4 ? atoi(argv[4]) : 20;
    int mode = argc > 5 ? atoi(argv[5]) : 0;
    if (argc > 6 && atoi(argv[6]) >= 0) { lrb_debug_set_scout(atoi(argv[6])); printf("scout tiles %d\n", atoi(argv[6])); }
    if (argc > 7) { lrb_debug_set_pair_drain(atoi(argv[7])); printf("pair drain %d\n", atoi(argv[7])); }
    if (argc > 8) { lrb_debug_set_overlap(atoi(argv[8])); printf("chunk overlap %d\n", atoi(argv[8])); }
    if (argc > 9) { lrb_debug_set_cap_div(atoi(argv[9])); printf("streams per user >= %d\n", atoi(argv[9])); }
    lrb_debug_set_score_mode(mode);
    if (mode) printf("debug mode %d\n", mode);
    const int which = argc > 10 ? atoi(argv[10]) : 3;   // bit 0: with bias + exclusion, bit 1: without
    if (argc > 11) { lrb_debug_set_restart_tiles(atof(argv[11])); printf("restart cost %s tiles\n", argv[11]); }
    if (which & 1) time_topk(B, rows, K, true, true);
    if (which & 2) time_topk(B, rows, K, false, false);
    return 0;
  }
  fails += check_dense(200, 1000);
  fails += check_dense(128, 256);
  fails += check_dense(1, 77);
  fails += check_topk(300, 50000, 20, 1, 0);
  fails += check_topk(300, 50000, 20, 0, 0);
  fails += check_topk(16, 1683, 20, 0, 0);
  fails += check_topk(700, 120000, 50, 0, 1000);
  fails += check_topk(130, 3000, 5, 0, 0);
  fails += check_topk(2048, 11001, 20, 0, 0);
  fails += check_topk(333, 70000, 20, 0, 0, false);
  fails += check_topk(4096, 300000, 20, 0, 0);
  printf(fails ? "FAILED (%d)\n" : "ALL OK\n", fails);
  return fails ? 1 : 0;
}
