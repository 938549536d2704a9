// LRURec encoder (north_star subsystem 1): embedding + LayerNorm, and per block
//   in_proj (complex, as a real 64->256 GEMM) * gamma      model/lru.py:151-153
//   complex-diagonal linear recurrence (exact tree-scan semantics)   model/lru.py:135-147,155-159
//   Re(out_proj) + residual + LayerNorm                    model/lru.py:160-161
//   PFFN: W2 gelu_erf(W1 y + b1) + b2 + y, LayerNorm       model/lru.py:164-175
//
// Tokens are processed in a COMPACT order: the positions before a user's first real item carry a
// False mask and sit before every real position, so they can never influence a later state
// (every hop out of them is multiplied by mask = 0) and, in eval mode, they are skipped.  Token i
// of the compact list belongs to user b = upper_bound(tok_offset, i) - 1 at position
// tok_first[b] + (i - tok_offset[b]).
//
// Three kernels per block, fp32 FFMA throughout (the 1e-3 score tolerance and the bit-identical
// candidate lists on small catalogues rule out bf16 tensor-core math here):
//   embed_inproj_kernel : [64-token tile] gather/LN (block 0) or load x, GEMM 64->256, * gamma
//   lru_scan_kernel     : one thread per (user, channel); sequential in time, coalesced over
//                         channels, reads bu once and writes h once (HBM-bound)
//   outproj_ffn_kernel  : [64-token tile] GEMM 256->64 + LN, GEMM 64->256 + GELU, GEMM 256->64 + LN
#include "api_util.h"
#include "common.cuh"

#include <climits>

namespace lrb {
namespace enc {

constexpr int D = LRB_D;        // 64
constexpr int H2 = 2 * LRB_H;   // 256 real numbers = 128 complex channels (re, im interleaved)
constexpr int FF = LRB_FF;      // 256
constexpr int TOK = 64;         // tokens per tile
constexpr int THREADS = 256;
constexpr float LN_EPS = 1e-5f;

// ---- packed weight blob (floats) -----------------------------------------------------------
constexpr int OFF_EMB_LN_W = 0;
constexpr int OFF_EMB_LN_B = 64;
constexpr int OFF_BLOCKS = 128;
constexpr int B_LAM_RE = 0;
constexpr int B_LAM_IM = 128;
constexpr int B_GAMMA = 256;
constexpr int B_WIN_T = 384;                     // [64][256]
constexpr int B_BIN = B_WIN_T + D * H2;          // [256]
constexpr int B_WOUT_T = B_BIN + H2;             // [256][64]
constexpr int B_BOUT = B_WOUT_T + H2 * D;        // [64]
constexpr int B_LN1_W = B_BOUT + D;
constexpr int B_LN1_B = B_LN1_W + D;
constexpr int B_W1_T = B_LN1_B + D;              // [64][256]
constexpr int B_B1 = B_W1_T + D * FF;            // [256]
constexpr int B_W2_T = B_B1 + FF;                // [256][64]
constexpr int B_B2 = B_W2_T + FF * D;            // [64]
constexpr int B_LN2_W = B_B2 + D;
constexpr int B_LN2_B = B_LN2_W + D;
constexpr int BLOCK_FLOATS = B_LN2_B + D;        // 66,816

// ---- token bookkeeping ----------------------------------------------------------------------
LRB_DEVINL int user_of_token(const int* __restrict__ tok_offset, int B, int i) {
  // largest b with tok_offset[b] <= i
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(tok_offset + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- shared-memory GEMM micro kernels -------------------------------------------------------
// A is stored k-major ("transposed"): At[k][TOK]; W is stored k-major: W[k][N].
// (A) K=64 -> N=256: thread owns 8 tokens x 8 columns (columns tx*4..+3 and 128+tx*4..+3).
LRB_DEVINL void gemm_k64_n256(const float* __restrict__ At, const float* __restrict__ W, int ty, int tx,
                              float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(At + k * TOK + ty * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(At + k * TOK + ty * 8 + 4);
    const float4 w0 = *reinterpret_cast<const float4*>(W + k * 256 + tx * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(W + k * 256 + 128 + tx * 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}
// (B) K=256 -> N=64: thread owns 4 tokens x 4 columns (ty = tid/16, tx = tid%16).
LRB_DEVINL void gemm_k256_n64(const float* __restrict__ At, const float* __restrict__ W, int ty, int tx,
                              float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int k = 0; k < 256; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(At + k * TOK + ty * 4);
    const float4 w4 = *reinterpret_cast<const float4*>(W + k * 64 + tx * 4);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

LRB_DEVINL void copy_weights(float* dst, const float* __restrict__ src, int n_floats, int tid) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < n_floats / 4; i += THREADS) d4[i] = __ldg(s4 + i);
}

// LayerNorm over 64 features of `rows` rows held in Cs[row][65]; one warp per row at a time.
// Two-pass (mean, then centred variance) like torch's CPU kernel.  The result replaces the row in
// Cs (row-major, conflict-free) and optionally goes to global memory (row-major [row][64]).
LRB_DEVINL void layernorm_rows(float* Cs, int n_rows, const float* __restrict__ w,
                               const float* __restrict__ b, float* dst_g_base,
                               const int* s_dst_row, int warp, int lane) {
  const float w0 = __ldg(w + lane), w1 = __ldg(w + lane + 32);
  const float b0 = __ldg(b + lane), b1 = __ldg(b + lane + 32);
  for (int r = warp; r < n_rows; r += THREADS / 32) {
    const float v0 = Cs[r * 65 + lane], v1 = Cs[r * 65 + lane + 32];
    const float mean = warp_sum(v0 + v1) * (1.0f / 64.0f);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float var = warp_sum(d0 * d0 + d1 * d1) * (1.0f / 64.0f);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
    const float o0 = d0 * rstd * w0 + b0;
    const float o1 = d1 * rstd * w1 + b1;
    Cs[r * 65 + lane] = o0;
    Cs[r * 65 + lane + 32] = o1;
    if (dst_g_base != nullptr) {
      const int row = s_dst_row[r];
      if (row >= 0) {
        dst_g_base[static_cast<size_t>(row) * D + lane] = o0;
        dst_g_base[static_cast<size_t>(row) * D + lane + 32] = o1;
      }
    }
  }
}

// Row-major Cs[row][65] -> k-major At[k][TOK]; lanes walk the rows, so both sides are conflict-free.
LRB_DEVINL void transpose_rows_to_kmajor(const float* Cs, float* At, int tid) {
  for (int e = tid; e < TOK * D; e += THREADS) {
    const int r = e & (TOK - 1), k = e / TOK;
    At[k * TOK + r] = Cs[r * 65 + k];
  }
}

// =============================================================================================
// Kernel 1: (embedding gather + LayerNorm | load x) -> in_proj -> * gamma -> bu[T][256]
// =============================================================================================
struct InprojParams {
  const void* ids;           // [B][L] int64 or int32
  int id_bytes;              // 8 or 4
  const float* table;        // [rows][64]
  long long table_rows;
  const float* wts;          // packed blob
  int blk;
  int first_block;           // 1: gather + LN from the table, also writes x0; 0: read x_in
  const float* x_in;         // [T][64] (blocks > 0)
  float* x0_out;             // [T][64] (block 0)
  float* bu;                 // [T][256]
  const int* tok_first;
  const int* tok_offset;     // [B+1]
  int B, L;
};

constexpr int SMEM_INPROJ = (D * TOK + D * H2 + TOK * 65) * 4 + TOK * 4;

// 97 KB of shared memory and <= 128 registers: two CTAs per SM (the id -> row gather is latency-bound and
// wants the extra warps in flight).
__global__ void __launch_bounds__(THREADS, 2) embed_inproj_kernel(const InprojParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* At = reinterpret_cast<float*>(smem_raw);   // [64][TOK]
  float* Ws = At + D * TOK;                          // [64][256]
  float* Cs = Ws + D * H2;                           // [TOK][65] staging for LN
  int* s_row = reinterpret_cast<int*>(Cs + TOK * 65);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* blkw = p.wts + OFF_BLOCKS + static_cast<size_t>(p.blk) * BLOCK_FLOATS;
  const int T = __ldg(p.tok_offset + p.B);
  const int n_tiles = (T + TOK - 1) / TOK;

  copy_weights(Ws, blkw + B_WIN_T, D * H2, tid);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int i0 = tile * TOK;
    __syncthreads();   // previous tile's readers of At / Cs are done
    if (p.first_block) {
      // gather 64 rows (256 B each) into Cs, one warp per row
      for (int r = warp; r < TOK; r += THREADS / 32) {
        const int i = i0 + r;
        float v0 = 0.f, v1 = 0.f;
        if (i < T) {
          const int b = user_of_token(p.tok_offset, p.B, i);
          const int t = __ldg(p.tok_first + b) + (i - __ldg(p.tok_offset + b));
          long long id = load_id(p.ids, static_cast<size_t>(b) * p.L + t, p.id_bytes);
          if (id < 0 || id >= p.table_rows) id = 0;   // out-of-range ids are clamped to the pad row
          const float* row = p.table + static_cast<size_t>(id) * D;
          v0 = __ldg(row + lane);
          v1 = __ldg(row + lane + 32);
        }
        Cs[r * 65 + lane] = v0;
        Cs[r * 65 + lane + 32] = v1;
        if (lane == 0) s_row[r] = i < T ? i : -1;
      }
      __syncthreads();
      layernorm_rows(Cs, TOK, p.wts + OFF_EMB_LN_W, p.wts + OFF_EMB_LN_B, p.x0_out, s_row, warp, lane);
      __syncthreads();
      transpose_rows_to_kmajor(Cs, At, tid);
    } else {
      // load x tile (row-major [i][64]) transposed into At; lanes walk the rows (conflict-free stores)
      for (int e = tid; e < TOK * D / 4; e += THREADS) {
        const int r = e & (TOK - 1);
        const int k4 = e / TOK;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i0 + r < T) v = *reinterpret_cast<const float4*>(p.x_in + static_cast<size_t>(i0 + r) * D + k4 * 4);
        At[(k4 * 4 + 0) * TOK + r] = v.x;
        At[(k4 * 4 + 1) * TOK + r] = v.y;
        At[(k4 * 4 + 2) * TOK + r] = v.z;
        At[(k4 * 4 + 3) * TOK + r] = v.w;
      }
    }
    __syncthreads();

    const int ty = tid >> 5, tx = tid & 31;
    float acc[8][8];
    gemm_k64_n256(At, Ws, ty, tx, acc);
    // epilogue: (+ b_in) * gamma, write bu (columns tx*4..+3 and 128+tx*4..+3)
    const float4 bi0 = __ldg(reinterpret_cast<const float4*>(blkw + B_BIN + tx * 4));
    const float4 bi1 = __ldg(reinterpret_cast<const float4*>(blkw + B_BIN + 128 + tx * 4));
    // column j belongs to complex channel j/2
    const float2 g0 = __ldg(reinterpret_cast<const float2*>(blkw + B_GAMMA + tx * 2));
    const float2 g1 = __ldg(reinterpret_cast<const float2*>(blkw + B_GAMMA + 64 + tx * 2));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i0 + ty * 8 + i;
      if (row < T) {
        float4 o0, o1;
        o0.x = (acc[i][0] + bi0.x) * g0.x; o0.y = (acc[i][1] + bi0.y) * g0.x;
        o0.z = (acc[i][2] + bi0.z) * g0.y; o0.w = (acc[i][3] + bi0.w) * g0.y;
        o1.x = (acc[i][4] + bi1.x) * g1.x; o1.y = (acc[i][5] + bi1.y) * g1.x;
        o1.z = (acc[i][6] + bi1.z) * g1.y; o1.w = (acc[i][7] + bi1.w) * g1.y;
        float* dst = p.bu + static_cast<size_t>(row) * H2;
        *reinterpret_cast<float4*>(dst + tx * 4) = o0;
        *reinterpret_cast<float4*>(dst + 128 + tx * 4) = o1;
      }
    }
  }
}

// =============================================================================================
// Kernel 2: the linear recurrence.  One CTA (128 threads = complex channels) per user.
//
// Exact restatement of the reference's recursive-doubling scan for ARBITRARY masks: with positions
// indexed in the power-of-two padded frame, p receives, for every set bit `lev` of p,
//     lambda^(p - q) * mask[q] * H_lev(q),   q = last position of the preceding aligned 2^lev block,
// where H_lev(q) is q's value after the levels below `lev` only.  Each thread keeps one running
// product per level (Q[lev], multiplied by lambda every step), i.e. a Fenwick-style carry set; for
// left-padded inputs this collapses to h_p = lambda * (mask[p-1] h_{p-1}) + bu_p.
// =============================================================================================
struct ScanParams {
  float2* bu;                // [T][128] complex, overwritten with h (unless last_only)
  const void* ids;           // [B][L] int64 or int32
  int id_bytes;
  const float* wts;
  int blk;
  const int* tok_first;
  const int* tok_offset;
  int B, L, levels;          // levels = log2(padded length)
  int last_only;             // 1: write only the state of the last token of each user to h_last
  float2* h_last;            // [B][128]
};

constexpr int MAX_LEVELS = 8;   // LRB_MAX_LEN = 256
constexpr int SCAN_CHUNK = 8;   // time steps whose loads are in flight together

__global__ void __launch_bounds__(128) lru_scan_kernel(const ScanParams p) {
  __shared__ unsigned char s_mask[LRB_MAX_LEN];
  const int b = blockIdx.x;
  const int c = threadIdx.x;
  const int first = __ldg(p.tok_first + b);
  const int base = __ldg(p.tok_offset + b);
  const int n = __ldg(p.tok_offset + b + 1) - base;
  const int off = (1 << p.levels) - p.L;   // left pad of the power-of-two frame
  for (int t = c; t < p.L; t += 128) s_mask[t] = load_id(p.ids, static_cast<size_t>(b) * p.L + t, p.id_bytes) > 0 ? 1 : 0;
  __syncthreads();
  const float* blkw = p.wts + OFF_BLOCKS + static_cast<size_t>(p.blk) * BLOCK_FLOATS;
  const float lr = __ldg(blkw + B_LAM_RE + c), li = __ldg(blkw + B_LAM_IM + c);
  float qr[MAX_LEVELS], qi[MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < MAX_LEVELS; ++l) { qr[l] = 0.f; qi[l] = 0.f; }

  // Time is walked in chunks of SCAN_CHUNK steps: the chunk's bu values are requested together (independent
  // 8-byte loads, one 1 KB row per step across the CTA), then the recurrence runs over registers and the states are
  // written back.  With a mean of ~9 real tokens per user the kernel is latency-bound, not bandwidth-bound: one
  // chunk puts SCAN_CHUNK rows in flight per thread instead of the single prefetched row of the first version.
  float2* row = p.bu + static_cast<size_t>(base) * LRB_H + c;

  // Fast path: a LEFT-PADDED row (no padding after a real item -- what the reference's dataloaders produce,
  // dataloader/lru.py:98-118,129-180).  The tree scan then equals  h_p = lambda * (mask_{p-1} h_{p-1}) + bu_p
  // (SURVEY probe P1): 4 FMAs per step instead of the ~150 instructions of the general carry set below, which
  // made this kernel instruction-bound (ncu round 2: issue active 71 %, 157 warp instructions per warp and step).
  bool monotone = true;
  for (int t = first + 1 + c; t < p.L; t += 128) monotone = monotone && !(s_mask[t - 1] != 0 && s_mask[t] == 0);
  if (__syncthreads_and(monotone ? 1 : 0)) {
    float hr = 0.f, hi = 0.f;   // mask_{p-1} * h_{p-1}
    for (int j0 = 0; j0 < n; j0 += SCAN_CHUNK) {
      float2 v[SCAN_CHUNK];
#pragma unroll
      for (int u = 0; u < SCAN_CHUNK; ++u)
        v[u] = (j0 + u < n) ? row[static_cast<size_t>(j0 + u) * LRB_H] : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < SCAN_CHUNK; ++u) {
        const int j = j0 + u;
        if (j < n) {
          const float nr = v[u].x + (hr * lr - hi * li);
          const float ni = v[u].y + (hr * li + hi * lr);
          if (!p.last_only) row[static_cast<size_t>(j) * LRB_H] = make_float2(nr, ni);
          else if (j == n - 1) p.h_last[static_cast<size_t>(b) * LRB_H + c] = make_float2(nr, ni);
          const bool m = s_mask[first + j] != 0;
          hr = m ? nr : 0.f;
          hi = m ? ni : 0.f;
        }
      }
    }
    return;
  }

  for (int j0 = 0; j0 < n; j0 += SCAN_CHUNK) {
    float2 v[SCAN_CHUNK];
#pragma unroll
    for (int u = 0; u < SCAN_CHUNK; ++u)
      v[u] = (j0 + u < n) ? row[static_cast<size_t>(j0 + u) * LRB_H] : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < SCAN_CHUNK; ++u) {
      const int j = j0 + u;
      if (j < n) {
        const int t = first + j;
        const int pos = t + off;
        float ar = v[u].x, ai = v[u].y;
        float sr = 0.f, si = 0.f;
        int zlev = -1;
        bool below_all_ones = true;
#pragma unroll
        for (int l = 0; l < MAX_LEVELS; ++l) {
          if (l < p.levels) {
            const bool bit = (pos >> l) & 1;
            if (bit) {
              ar += qr[l];
              ai += qi[l];
            } else if (below_all_ones) {
              sr = ar; si = ai; zlev = l;     // H_l(pos): value after the levels below l only
              below_all_ones = false;
            }
          }
        }
        if (!p.last_only) row[static_cast<size_t>(j) * LRB_H] = make_float2(ar, ai);
        else if (j == n - 1) p.h_last[static_cast<size_t>(b) * LRB_H + c] = make_float2(ar, ai);
        const float m = s_mask[t] ? 1.f : 0.f;
#pragma unroll
        for (int l = 0; l < MAX_LEVELS; ++l) {
          if (l < p.levels) {
            float xr = qr[l], xi = qi[l];
            if (l == zlev) { xr = sr * m; xi = si * m; }
            qr[l] = xr * lr - xi * li;
            qi[l] = xr * li + xi * lr;
          }
        }
      }
    }
  }
}

// =============================================================================================
// Kernel 3: Re(out_proj h) + x -> LN -> PFFN -> LN
// =============================================================================================
struct OutParams {
  const float* h;            // [T][256] (all-token mode) or h_last [B][256] (last-only mode)
  const float* x_res;        // [T][64] residual input of this block
  const float* wts;
  int blk;
  float* x_out;              // [T][64] or [B][64] (last-only)
  __nv_bfloat16* out_bf16;   // optional bf16 copy (last-only mode)
  const int* tok_offset;
  int B;
  int last_only;
};

constexpr int SMEM_OUT = (H2 * TOK + H2 * D + D * TOK + TOK * 65 + TOK * 65) * 4 + TOK * 4;

__global__ void __launch_bounds__(THREADS, 1) outproj_ffn_kernel(const OutParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* Ht = reinterpret_cast<float*>(smem_raw);   // [256][TOK]   h, later the PFFN hidden layer
  float* Ws = Ht + H2 * TOK;                         // 64 KB weight stage
  float* Yt = Ws + H2 * D;                           // [64][TOK]    y (k-major) for the W1 GEMM
  float* Cs = Yt + D * TOK;                          // [TOK][65]    LN staging
  float* Ys = Cs + TOK * 65;                         // [TOK][65]    y row-major (PFFN residual)
  int* s_row = reinterpret_cast<int*>(Ys + TOK * 65);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* blkw = p.wts + OFF_BLOCKS + static_cast<size_t>(p.blk) * BLOCK_FLOATS;
  const int T = __ldg(p.tok_offset + p.B);
  const int n_rows_total = p.last_only ? p.B : T;
  const int n_tiles = (n_rows_total + TOK - 1) / TOK;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int i0 = tile * TOK;
    __syncthreads();
    // ---- stage h (transposed), the residual rows, W_out ----
    for (int e = tid; e < TOK * H2 / 4; e += THREADS) {
      const int r = e & (TOK - 1);      // lanes walk the rows: conflict-free transposed stores
      const int k4 = e / TOK;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + r < n_rows_total) v = *reinterpret_cast<const float4*>(p.h + static_cast<size_t>(i0 + r) * H2 + k4 * 4);
      Ht[(k4 * 4 + 0) * TOK + r] = v.x;
      Ht[(k4 * 4 + 1) * TOK + r] = v.y;
      Ht[(k4 * 4 + 2) * TOK + r] = v.z;
      Ht[(k4 * 4 + 3) * TOK + r] = v.w;
    }
    for (int e = tid; e < TOK * D; e += THREADS) {
      const int r = e >> 6, k = e & 63;
      float v = 0.f;
      const int i = i0 + r;
      if (i < n_rows_total) {
        // last-only mode: row i is user i, its residual is the block input at that user's last token
        const int src = p.last_only ? (__ldg(p.tok_offset + i + 1) - 1) : i;
        v = p.x_res[static_cast<size_t>(src) * D + k];
      }
      Cs[r * 65 + k] = v;
    }
    if (tid < TOK) s_row[tid] = (i0 + tid < n_rows_total) ? i0 + tid : -1;
    copy_weights(Ws, blkw + B_WOUT_T, H2 * D, tid);
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    {
      // ---- y = LN(h W_out^T(real form) + b_out + x) ----
      float acc[4][4];
      gemm_k256_n64(Ht, Ws, ty, tx, acc);
      const float4 bo = __ldg(reinterpret_cast<const float4*>(blkw + B_BOUT + tx * 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        Cs[r * 65 + tx * 4 + 0] += acc[i][0] + bo.x;
        Cs[r * 65 + tx * 4 + 1] += acc[i][1] + bo.y;
        Cs[r * 65 + tx * 4 + 2] += acc[i][2] + bo.z;
        Cs[r * 65 + tx * 4 + 3] += acc[i][3] + bo.w;
      }
    }
    __syncthreads();
    // LN in place (row-major y, also the PFFN residual), then a conflict-free transpose into Yt
    layernorm_rows(Cs, TOK, blkw + B_LN1_W, blkw + B_LN1_B, nullptr, s_row, warp, lane);
    copy_weights(Ws, blkw + B_W1_T, D * FF, tid);   // W_out no longer needed (all GEMM reads done)
    __syncthreads();
    for (int e = tid; e < TOK * D; e += THREADS) {
      const int r = e & (TOK - 1), k = e / TOK;
      const float yv = Cs[r * 65 + k];
      Yt[k * TOK + r] = yv;
      Ys[r * 65 + k] = yv;
    }
    __syncthreads();
    {
      // ---- f = gelu_erf(y W1^T + b1) -> Ht (k-major [256][TOK]) ----
      const int ty8 = tid >> 5, tx8 = tid & 31;
      float acc[8][8];
      gemm_k64_n256(Yt, Ws, ty8, tx8, acc);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(blkw + B_B1 + tx8 * 4));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(blkw + B_B1 + 128 + tx8 * 4));
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      __syncthreads();   // every warp finished reading Ht? (not read here) -- keeps Ws/Yt readers aligned
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = (j < 4 ? tx8 * 4 + j : 128 + tx8 * 4 + (j - 4));
        float4 o0, o1;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = acc[i][j] + bb[j];
          v[i] = 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
        }
        o0 = make_float4(v[0], v[1], v[2], v[3]);
        o1 = make_float4(v[4], v[5], v[6], v[7]);
        *reinterpret_cast<float4*>(Ht + col * TOK + ty8 * 8) = o0;
        *reinterpret_cast<float4*>(Ht + col * TOK + ty8 * 8 + 4) = o1;
      }
    }
    __syncthreads();
    copy_weights(Ws, blkw + B_W2_T, FF * D, tid);
    __syncthreads();
    {
      // ---- out = LN(f W2^T + b2 + y) ----
      float acc[4][4];
      gemm_k256_n64(Ht, Ws, ty, tx, acc);
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(blkw + B_B2 + tx * 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        Cs[r * 65 + tx * 4 + 0] = acc[i][0] + b2.x + Ys[r * 65 + tx * 4 + 0];
        Cs[r * 65 + tx * 4 + 1] = acc[i][1] + b2.y + Ys[r * 65 + tx * 4 + 1];
        Cs[r * 65 + tx * 4 + 2] = acc[i][2] + b2.z + Ys[r * 65 + tx * 4 + 2];
        Cs[r * 65 + tx * 4 + 3] = acc[i][3] + b2.w + Ys[r * 65 + tx * 4 + 3];
      }
    }
    __syncthreads();
    layernorm_rows(Cs, TOK, blkw + B_LN2_W, blkw + B_LN2_B, p.x_out, s_row, warp, lane);
    if (p.out_bf16 != nullptr) {
      // bf16 copy of the rows this warp just normalised in place (same warp -> same rows, no barrier)
      for (int r = warp; r < TOK; r += THREADS / 32) {
        const int rowi = s_row[r];
        if (rowi < 0) continue;
        p.out_bf16[static_cast<size_t>(rowi) * D + lane] = __float2bfloat16_rn(Cs[r * 65 + lane]);
        p.out_bf16[static_cast<size_t>(rowi) * D + lane + 32] = __float2bfloat16_rn(Cs[r * 65 + lane + 32]);
      }
    }
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace enc
}  // namespace lrb

extern "C" {

size_t lrb_encoder_weight_floats(int n_blocks) {
  return static_cast<size_t>(lrb::enc::OFF_BLOCKS) + static_cast<size_t>(n_blocks) * lrb::enc::BLOCK_FLOATS;
}

size_t lrb_encode_workspace_bytes(int B, int L, int all_positions) {
  using namespace lrb::enc;
  (void)all_positions;
  const size_t T = static_cast<size_t>(B) * L;
  // xa, xb: [T][64] ping-pong block inputs/outputs; bu: [T][256]; h_last: [B][256]
  return 2 * align256(T * D * 4) + align256(T * H2 * 4) + align256(static_cast<size_t>(B) * H2 * 4);
}

int lrb_encode_fwd(const void* ids, int id_bytes, int B, int L, const float* table_f32, int64_t table_rows,
                   const float* weights, int n_blocks, int all_positions, const int32_t* tok_first,
                   const int32_t* tok_offset, float* out_f32, void* out_bf16, void* workspace,
                   size_t workspace_bytes, void* stream) {
  using namespace lrb;
  using namespace lrb::enc;
  int rc = check_arch();
  if (rc != LRB_OK) return rc;
  LRB_REQUIRE(ids && table_f32 && weights && tok_first && tok_offset && out_f32 && workspace,
              "lrb_encode_fwd: null pointer");
  LRB_REQUIRE(B > 0 && L > 0 && n_blocks >= 1 && table_rows > 0, "lrb_encode_fwd: bad shape");
  LRB_REQUIRE(id_bytes == 4 || id_bytes == 8, "lrb_encode_fwd: id_bytes must be 4 (int32) or 8 (int64)");
  if (L > LRB_MAX_LEN)
    return set_error(LRB_ERR_UNSUPPORTED, "sequence length %d exceeds LRB_MAX_LEN=%d", L, LRB_MAX_LEN);
  if (workspace_bytes < lrb_encode_workspace_bytes(B, L, all_positions))
    return set_error(LRB_ERR_WORKSPACE, "lrb_encode_fwd: workspace too small");
  LRB_REQUIRE(!(all_positions && out_bf16 != nullptr), "lrb_encode_fwd: out_bf16 is an eval-mode output");
  cudaStream_t st = as_stream(stream);
  const size_t T = static_cast<size_t>(B) * L;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* xa = reinterpret_cast<float*>(ws);
  float* xb = reinterpret_cast<float*>(ws + align256(T * D * 4));
  float* bu = reinterpret_cast<float*>(ws + 2 * align256(T * D * 4));
  float* h_last = reinterpret_cast<float*>(ws + 2 * align256(T * D * 4) + align256(T * H2 * 4));
  int levels = 0;
  while ((1 << levels) < L) ++levels;

  LRB_CUDA_TRY(cudaFuncSetAttribute(embed_inproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_INPROJ));
  LRB_CUDA_TRY(cudaFuncSetAttribute(outproj_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OUT));
  const int sms = device_sm_count();
  const int max_tiles = static_cast<int>((T + TOK - 1) / TOK);
  const int grid_tok = max_tiles < 2 * sms ? max_tiles : 2 * sms;   // embed_inproj: two CTAs per SM

  float* x_cur = xa;    // output of the previous block / embedding
  float* x_nxt = xb;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const bool last_blk = blk == n_blocks - 1;
    const bool last_only = last_blk && !all_positions;
    InprojParams ip;
    ip.ids = ids; ip.id_bytes = id_bytes;
    ip.table = table_f32; ip.table_rows = table_rows; ip.wts = weights; ip.blk = blk;
    ip.first_block = blk == 0 ? 1 : 0;
    ip.x_in = x_cur; ip.x0_out = blk == 0 ? x_cur : nullptr; ip.bu = bu;
    ip.tok_first = tok_first; ip.tok_offset = tok_offset; ip.B = B; ip.L = L;
    embed_inproj_kernel<<<grid_tok, THREADS, SMEM_INPROJ, st>>>(ip);
    LRB_CUDA_TRY(cudaGetLastError());

    ScanParams sp;
    sp.bu = reinterpret_cast<float2*>(bu);
    sp.ids = ids; sp.id_bytes = id_bytes;
    sp.wts = weights; sp.blk = blk; sp.tok_first = tok_first; sp.tok_offset = tok_offset;
    sp.B = B; sp.L = L; sp.levels = levels; sp.last_only = last_only ? 1 : 0;
    sp.h_last = reinterpret_cast<float2*>(h_last);
    lru_scan_kernel<<<B, 128, 0, st>>>(sp);
    LRB_CUDA_TRY(cudaGetLastError());

    OutParams op;
    op.h = last_only ? h_last : bu;
    op.x_res = x_cur; op.wts = weights; op.blk = blk;
    op.x_out = last_blk ? out_f32 : x_nxt;
    op.out_bf16 = last_only ? static_cast<__nv_bfloat16*>(out_bf16) : nullptr;
    op.tok_offset = tok_offset; op.B = B; op.last_only = last_only ? 1 : 0;
    const int tiles = last_only ? (B + TOK - 1) / TOK : max_tiles;
    const int grid = tiles < sms ? tiles : sms;
    outproj_ffn_kernel<<<grid, THREADS, SMEM_OUT, st>>>(op);
    LRB_CUDA_TRY(cudaGetLastError());
    float* tmp = x_cur; x_cur = x_nxt; x_nxt = tmp;
  }
  return LRB_OK;
}

}  // extern "C"
