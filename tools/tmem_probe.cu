// Developer micro-benchmark (not part of the product path): how fast can the accumulators of a K = 64 GEMM be read
// out of TMEM while the next MMAs run?  DESIGN.md section 4.1 / 9: ~440 of the ~1260 cycles the scoring kernel
// spends per 128 x 256 tile are tcgen05.ld time that does not overlap the MMA.  This probe times, per SM,
//   mode 0: MMAs only (4 x M128 N256 K16 per tile, operands resident in shared memory)
//   mode 1: + every accumulator read as fp32   (4 x tcgen05.ld.32x32b.x32 per epilogue warp)
//   mode 2: + fp16 accumulators read packed    (2 x tcgen05.ld.32x32b.x32.pack::16b per epilogue warp)
//   mode 3: loads of mode 1 without MMAs, mode 4: loads of mode 2 without MMAs
//   mode 5: mode 1 + the scoring kernel's filter per 32 columns (FMNMX3 max tree + one compare), 8 epilogue warps
//   mode 6: the same with 16 epilogue warps (4 per SM sub-partition, 64 columns per thread)
//   mode 7: 8 epilogue warps, loads double-buffered in PAIRS of 32-column chunks (128 load registers per thread,
//           setmaxnreg: producer/MMA warpgroup down to 40 registers, epilogue warpgroups up to 232)
// Measured on B200 (round 1, 3000 tiles per SM, 1965 MHz): mode 0 = 520, mode 1 = 623 cycles per tile; mode 2 raised
// "illegal instruction" (f16 accumulators and/or .pack::16b in this form) -- it only runs when asked for (argv[2]).
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I llamarec_b200/csrc -o tools/tmem_probe tools/tmem_probe.cu
// run:    tools/tmem_probe [tiles per SM]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace lrb;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int ACC = 2;

LRB_DEVINL float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

LRB_DEVINL void tmem_ld_32x32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int EPI_WARPS, bool PAIRBUF>
__global__ void __launch_bounds__(128 + EPI_WARPS * 32, 1)
probe_kernel(int tiles, int mode, long long* cycles, unsigned* sink, float thr, int random_data, int n128) {
  constexpr int THREADS = 128 + EPI_WARPS * 32;
  constexpr int PARTS = EPI_WARPS / 4;          // column parts per lane quadrant
  constexpr int COLS = BN / PARTS;              // columns per epilogue thread
  constexpr int CHUNKS = COLS / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                        // [128][64] bf16, 128-byte swizzle atoms (contents irrelevant)
  uint8_t* sB = smem + BM * BK * 2;          // [256][64] bf16
  __shared__ uint64_t full_bar[ACC], empty_bar[ACC];
  __shared__ uint64_t sfull[4], sempty[4];      // modes 8-10: the scoring kernel's shared-memory stage handshake
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mma_on = mode <= 2 || mode >= 5;
  const int extra = mode >= 8 ? mode - 7 : 0;   // 1: second commit per tile (nobody waits), 2: + wait on a stage barrier
                                                // re-armed by a producer thread, 3: the same with clock64 reads around the waits
  if (mode >= 8) mode = 0;
  const int ld_mode = mode == 0 ? 0 : (PAIRBUF ? 3 : ((mode == 1 || mode == 3 || mode >= 5) ? 1 : 2));
  const bool filter = mode >= 5;
  for (int i = threadIdx.x; i < (BM + BN) * BK * 2 / 4; i += THREADS) {
    uint32_t v = 0x3c003c00u;
    if (random_data) {   // pseudo-random bf16 pairs in (-2, 2): realistic operand toggling (tensor-core power)
      uint32_t h = (i + 1u) * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x807f807fu) | 0x3f003f00u | ((h >> 3) & 0x00800080u);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < ACC; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], EPI_WARPS); }
    for (int i = 0; i < 4; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 1); }
    mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(&tmem_ptr, ACC * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (PAIRBUF) {   // uniform per warpgroup (warps 0-3 | 4-7 | 8-11)
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  }
  const long long t0 = clock64();
  if (warp == 1 && lane == 0) {
    uint32_t idesc = umma_idesc_bf16(BM, n128 ? 128 : BN);
    if (ld_mode == 2) idesc &= ~(3u << 4);                 // c_format = F16 accumulators
    const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sB));
    int acc = 0; uint32_t ph = 0;
    int stg = 0; uint32_t sph = 0;
    long long wsum = 0;
    for (int t = 0; t < tiles; ++t) {
      if (extra >= 2) {
        const long long w0 = extra >= 3 ? clock64() : 0;
        mbar_wait(&sfull[stg], sph);
        if (extra >= 3) wsum += clock64() - w0;
      }
      mbar_wait(&empty_bar[acc], ph ^ 1);
      tc_fence_after();
      if (mma_on) {
        if (n128) {   // the same tile as two N = 128 halves (8 MMAs)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(tmem_base + acc * BN + hh * 128, da + 2 * k, db + hh * 1024 + 2 * k, idesc, k > 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16_ss(tmem_base + acc * BN, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
        }
      }
      umma_commit(&full_bar[acc]);
      if (extra >= 1) {
        umma_commit(&sempty[stg]);
        if (++stg == 3) { stg = 0; sph ^= 1; }
      }
      if (++acc == ACC) { acc = 0; ph ^= 1; }
    }
    if (wsum == 0x7fffffff) sink[0] = 1;
  } else if (warp == 0 && lane == 0 && extra >= 2) {
    int stg = 0; uint32_t sph = 0;
    for (int t = 0; t < tiles; ++t) {      // "producer": stage free -> hand it over again
      mbar_wait(&sempty[stg], sph ^ 1);
      mbar_arrive(&sfull[stg]);
      if (++stg == 3) { stg = 0; sph ^= 1; }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, quad = ew & 3, half = ew >> 2;
    int acc = 0; uint32_t ph = 0;
    unsigned x = 0;
    int hits = 0;
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(&full_bar[acc], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN + half * COLS);
      if (ld_mode == 1) {
        uint32_t v[2][32];
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          if (c < CHUNKS - 1) tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
          if (filter) {
            // the scoring kernel's hot path: group maxima of 2 x 16 scores, one compare per 32 columns
            float gm[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float q[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) q[j] = __uint_as_float(v[c & 1][g * 16 + j]);
              gm[g] = fmaxf(max3f(max3f(q[0], q[1], q[2]), max3f(q[3], q[4], q[5]), max3f(q[6], q[7], q[8])),
                            max3f(max3f(q[9], q[10], q[11]), max3f(q[12], q[13], q[14]), q[15]));
            }
            if (fmaxf(gm[0], gm[1]) >= thr) ++hits;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x ^= v[c & 1][j];
          }
          if (c < CHUNKS - 1) tmem_ld_wait();
        }
      } else if (PAIRBUF && ld_mode == 3) {
        // 4 chunks as two pairs: the second pair is in flight while the first is filtered
        uint32_t a0[32], a1[32], b0[32], b1[32];
        tmem_ld_32x32(taddr, a0);
        tmem_ld_32x32(taddr + 32, a1);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 64, b0);
        tmem_ld_32x32(taddr + 96, b1);
        auto filt = [&](const uint32_t (&v)[32]) {
          float gm[2];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) q[j] = __uint_as_float(v[g * 16 + j]);
            gm[g] = fmaxf(max3f(max3f(q[0], q[1], q[2]), max3f(q[3], q[4], q[5]), max3f(q[6], q[7], q[8])),
                          max3f(max3f(q[9], q[10], q[11]), max3f(q[12], q[13], q[14]), q[15]));
          }
          if (fmaxf(gm[0], gm[1]) >= thr) ++hits;
        };
        filt(a0);
        filt(a1);
        tmem_ld_wait();
        filt(b0);
        filt(b1);
      } else if (ld_mode == 2) {
        uint32_t v[2][32];
        tmem_ld_32x32_pack16(taddr, v[0]);           // 64 columns -> 32 registers
        tmem_ld_32x32_pack16(taddr + 64, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) x ^= v[0][j] ^ v[1][j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[acc]);
      if (++acc == ACC) { acc = 0; ph ^= 1; }
    }
    if (x == 0x12345678u || hits == 0x7fffffff) sink[0] = x + hits;
    // (taken by a thread that really waited for the last tile: BAR.SYNC is deferred-blocking, a clock read right
    // after __syncthreads() by an idle warp returns the time the barrier was ISSUED)
    if (threadIdx.x == 128) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, ACC * BN); }
}

static int g_random = 0, g_n128 = 0;
template <int EPI_WARPS, bool PAIRBUF = false>
static double run(int sms, int tiles, int mode, long long* d_cycles, unsigned* d_sink) {
  const int smem = (BM + BN) * BK * 2 + 1024;
  cudaFuncSetAttribute(probe_kernel<EPI_WARPS, PAIRBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    probe_kernel<EPI_WARPS, PAIRBUF><<<sms, 128 + EPI_WARPS * 32, smem>>>(tiles, mode, d_cycles, d_sink, 1e30f, g_random, g_n128);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long c : h) mean += c;
  return mean / sms / tiles;
}

int main(int argc, char** argv) {
  const int tiles = argc > 1 ? atoi(argv[1]) : 4000;
  const bool try_f16 = false;
  g_random = argc > 2 ? atoi(argv[2]) : 0;
  g_n128 = argc > 3 ? atoi(argv[3]) : 0;
  printf("operands: %s, MMA N = %d\n", g_random ? "pseudo-random" : "constant", g_n128 ? 128 : 256);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cycles; unsigned* d_sink;
  cudaMalloc(&d_cycles, sms * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  for (int m = 8; m <= 10; ++m)
    printf("mode %d (MMA only + %s): %.0f cycles per tile per SM\n", m,
           m == 8 ? "a second tcgen05.commit per tile" : m == 9 ? "second commit + stage-barrier wait (producer thread re-arms)"
                                                                 : "second commit + stage-barrier wait + clock64 reads",
           run<8>(sms, tiles, m, d_cycles, d_sink));
  const char* names[8] = {"MMA only", "MMA + fp32 read-out", "MMA(f16 acc) + packed 16-bit read-out", "fp32 read-out only",
                          "packed 16-bit read-out only", "MMA + fp32 read-out + max-tree filter, 8 epilogue warps",
                          "MMA + fp32 read-out + max-tree filter, 16 epilogue warps",
                          "MMA + fp32 read-out + max-tree filter, 8 warps, pair-buffered loads (setmaxnreg)"};
  const int order[8] = {0, 1, 3, 5, 6, 7, 2, 4};
  for (int i = 0; i < 8; ++i) {
    const int mode = order[i];
    if ((mode == 2 || mode == 4) && !try_f16) continue;
    const double cyc = mode == 6 ? run<16>(sms, tiles, mode, d_cycles, d_sink)
                     : mode == 7 ? run<8, true>(sms, tiles, mode, d_cycles, d_sink)
                                 : run<8>(sms, tiles, mode, d_cycles, d_sink);
    printf("mode %d (%s): %.0f cycles per 128x256x64 tile per SM\n", mode, names[mode], cyc);
  }
  return 0;
}
