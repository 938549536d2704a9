// Developer micro-benchmark (not part of the product path): which epilogue structure reads a 128 x 256 fp32
// accumulator tile out of TMEM and runs the scoring kernel's max-tree threshold filter closest to the
// hardware floor (MMA 520 cycles, TMEM read-out 512 cycles per tile, tools/tmem_probe)?
// Same frame as tools/tmem_probe (operands resident in shared memory, two accumulator stages, one MMA thread),
// but the epilogue variants mimic the real kernel's per-tile bookkeeping (threshold read from shared memory,
// early release of the accumulator stage, one vote per tile, rare ring store).
//   variant 0: 8 warps, 128 columns / thread, x32 loads double-buffered one chunk ahead   (round-1 kernel)
//   variant 1: 16 warps, 64 columns / thread, x32 loads double-buffered one chunk ahead
//   variant 2: 16 warps, 64 columns / thread, both x32 loads issued up front, stage released before the filter
//   variant 3: 16 warps, 64 columns / thread, single x32 buffer, no prefetch
//   variant 4: 8 warps, 128 columns / thread, loads in pairs of chunks (128 load registers, setmaxnreg)
//   variant 5: 8 warps, 128 columns / thread, all four x32 loads up front (128 registers, setmaxnreg), stage
//              released before the filter
//   variant 6: 16 warps, 64 columns / thread, four x16 loads up front
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I llamarec_b200/csrc -o tools/epi_probe tools/epi_probe.cu
// run:    tools/epi_probe [tiles per SM]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace lrb;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int ACC = 2;

LRB_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

LRB_DEVINL float gmax16(const uint32_t* v) {
  float q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) q[j] = __uint_as_float(v[j]);
  return fmaxf(max3(max3(q[0], q[1], q[2]), max3(q[3], q[4], q[5]), max3(q[6], q[7], q[8])),
               max3(max3(q[9], q[10], q[11]), max3(q[12], q[13], q[14]), q[15]));
}

// the hot path of the scoring kernel for 32 columns: two group maxima, one branch; rare: dump to the ring
LRB_DEVINL void filter32(const uint32_t (&v)[32], float thr, float4* ring, int& cnt, int gid0) {
  const float g0 = gmax16(v), g1 = gmax16(v + 16);
  if (fmaxf(g0, g1) >= thr) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if ((g == 0 ? g0 : g1) >= thr) {
        float4* rec = ring + (cnt & 15) * 5;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          rec[j] = make_float4(__uint_as_float(v[g * 16 + 4 * j]), __uint_as_float(v[g * 16 + 4 * j + 1]),
                               __uint_as_float(v[g * 16 + 4 * j + 2]), __uint_as_float(v[g * 16 + 4 * j + 3]));
        reinterpret_cast<int*>(rec + 4)[0] = gid0 + g * 16;
        ++cnt;
      }
    }
  }
}
LRB_DEVINL void filter16(const uint32_t (&v)[16], float thr, float4* ring, int& cnt, int gid0) {
  const float g0 = gmax16(v);
  if (g0 >= thr) {
    float4* rec = ring + (cnt & 15) * 5;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      rec[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                           __uint_as_float(v[4 * j + 3]));
    reinterpret_cast<int*>(rec + 4)[0] = gid0;
    ++cnt;
  }
}

template <int EPI_WARPS, int VARIANT, int NONEPI_REGS, int EPI_REGS>
__global__ void __launch_bounds__(128 + EPI_WARPS * 32, 1)
probe_kernel(int tiles, long long* cycles, unsigned* sink, float thr_in, float4* ring_base) {
  constexpr int THREADS = 128 + EPI_WARPS * 32;
  constexpr int PARTS = EPI_WARPS / 4;          // column parts per lane quadrant
  constexpr int COLS = BN / PARTS;              // columns per epilogue thread
  constexpr int CHUNKS = COLS / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                        // [128][64] bf16, 128-byte swizzle atoms (contents irrelevant)
  uint8_t* sB = smem + BM * BK * 2;          // [256][64] bf16
  __shared__ uint64_t full_bar[ACC], empty_bar[ACC];
  __shared__ uint32_t tmem_ptr;
  __shared__ int sRowThr[BM];
  __shared__ int sDrainSeq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (BM + BN) * BK * 2 / 4; i += THREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < BM) sRowThr[threadIdx.x] = float_to_key(thr_in);
  if (threadIdx.x == 0) {
    sDrainSeq = 0;
    for (int i = 0; i < ACC; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(&tmem_ptr, ACC * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (NONEPI_REGS > 0) {   // uniform per warpgroup
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NONEPI_REGS > 0 ? NONEPI_REGS : 24));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS > 0 ? EPI_REGS : 24));
  }
  const long long t0 = clock64();
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(BM, BN);
    const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sB));
    int acc = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(&empty_bar[acc], ph ^ 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) umma_bf16_ss(tmem_base + acc * BN, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
      umma_commit(&full_bar[acc]);
      if (++acc == ACC) { acc = 0; ph ^= 1; }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, quad = ew & 3, part = ew >> 2;
    const int r = quad * 32 + lane;
    const uint32_t thr_addr = smem_u32(&sRowThr[r]);
    const uint32_t seq_addr = smem_u32(&sDrainSeq);
    float4* ring = ring_base + (static_cast<size_t>(blockIdx.x) * EPI_WARPS * 32 + ew * 32 + lane) * 16 * 5;
    int acc = 0; uint32_t ph = 0;
    int cnt = 0;
    int drain_seen = 0;
    float own_thr = -INFINITY;
    for (int t = 0; t < tiles; ++t) {
      // threshold of this row, read before the wait (ld.volatile.shared: another warp may raise it)
      int kk;
      asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(kk) : "r"(thr_addr));
      const float thr = fmaxf(own_thr, key_to_float(kk));
      mbar_wait(&full_bar[acc], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN + part * COLS);
      const int gid0 = t * BN + part * COLS;
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[acc]);
      };
      if (VARIANT == 0 || VARIANT == 1) {
        uint32_t v[2][32];
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          if (c < CHUNKS - 1) tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
          filter32(v[c & 1], thr, ring, cnt, gid0 + c * 32);
          if (c < CHUNKS - 1) tmem_ld_wait();
          if (c == CHUNKS - 2) release();
        }
      } else if (VARIANT == 2) {
        static_assert(VARIANT != 2 || CHUNKS == 2, "variant 2 is the 16-warp layout");
        uint32_t a[32], b[32];
        tmem_ld_32x32(taddr, a);
        tmem_ld_32x32(taddr + 32, b);
        tmem_ld_wait();
        release();
        filter32(a, thr, ring, cnt, gid0);
        filter32(b, thr, ring, cnt, gid0 + 32);
      } else if (VARIANT == 3) {
        uint32_t a[32];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          tmem_ld_32x32(taddr + c * 32, a);
          tmem_ld_wait();
          if (c == CHUNKS - 1) release();
          filter32(a, thr, ring, cnt, gid0 + c * 32);
        }
      } else if (VARIANT == 4) {
        uint32_t a0[32], a1[32], b0[32], b1[32];
        tmem_ld_32x32(taddr, a0);
        tmem_ld_32x32(taddr + 32, a1);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 64, b0);
        tmem_ld_32x32(taddr + 96, b1);
        filter32(a0, thr, ring, cnt, gid0);
        filter32(a1, thr, ring, cnt, gid0 + 32);
        tmem_ld_wait();
        release();
        filter32(b0, thr, ring, cnt, gid0 + 64);
        filter32(b1, thr, ring, cnt, gid0 + 96);
      } else if (VARIANT == 5) {
        uint32_t a0[32], a1[32], b0[32], b1[32];
        tmem_ld_32x32(taddr, a0);
        tmem_ld_32x32(taddr + 32, a1);
        tmem_ld_32x32(taddr + 64, b0);
        tmem_ld_32x32(taddr + 96, b1);
        tmem_ld_wait();
        release();
        filter32(a0, thr, ring, cnt, gid0);
        filter32(a1, thr, ring, cnt, gid0 + 32);
        filter32(b0, thr, ring, cnt, gid0 + 64);
        filter32(b1, thr, ring, cnt, gid0 + 96);
      } else if (VARIANT == 6) {
        uint32_t a[16], b[16], c[16], d[16];
        tmem_ld_32x16(taddr, a);
        tmem_ld_32x16(taddr + 16, b);
        tmem_ld_32x16(taddr + 32, c);
        tmem_ld_32x16(taddr + 48, d);
        tmem_ld_wait();
        release();
        filter16(a, thr, ring, cnt, gid0);
        filter16(b, thr, ring, cnt, gid0 + 16);
        filter16(c, thr, ring, cnt, gid0 + 32);
        filter16(d, thr, ring, cnt, gid0 + 48);
      }
      // once per tile: does any lane's ring need draining / did another warp ask for a CTA-wide drain?
      const bool need = __any_sync(0xffffffffu, cnt > 8);
      int seq;
      asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(seq) : "r"(seq_addr));
      if (need || seq != drain_seen) {
        drain_seen = seq;
        own_thr = fmaxf(own_thr, __uint_as_float(0x7f000000u));   // never taken with thr = 1e30
        cnt = 0;
      }
      if (++acc == ACC) { acc = 0; ph ^= 1; }
    }
    if (cnt == 0x7fffffff) sink[0] = cnt;
    // (taken by a thread that really waited for the last tile: BAR.SYNC is deferred-blocking, a clock read right
    // after __syncthreads() by an idle warp returns the time the barrier was ISSUED)
    if (threadIdx.x == 128) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, ACC * BN); }
}

template <int EPI_WARPS, int VARIANT, int NONEPI_REGS = 0, int EPI_REGS = 0>
static double run(int sms, int tiles, long long* d_cycles, unsigned* d_sink, float4* d_ring, const char* name) {
  const int smem = (BM + BN) * BK * 2 + 1024;
  auto kern = probe_kernel<EPI_WARPS, VARIANT, NONEPI_REGS, EPI_REGS>;
  cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ea != cudaSuccess) { printf("variant %d: set attribute: %s\n", VARIANT, cudaGetErrorString(ea)); exit(1); }
  cudaMemset(d_cycles, 0, sms * sizeof(long long));
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<sms, 128 + EPI_WARPS * 32, smem>>>(tiles, d_cycles, d_sink, 1e30f, d_ring);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", VARIANT, cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long c : h) mean += c;
  const double cyc = mean / sms / tiles;
  printf("variant %d (%s; %d regs, %zu B local): %.0f cycles per 128x256x64 tile per SM\n", VARIANT, name, fa.numRegs,
         (size_t)fa.localSizeBytes, cyc);
  fflush(stdout);
  return cyc;
}

int main(int argc, char** argv) {
  const int tiles = argc > 1 ? atoi(argv[1]) : 4000;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cycles; unsigned* d_sink; float4* d_ring;
  cudaMalloc(&d_cycles, sms * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  if (cudaMalloc(&d_ring, static_cast<size_t>(sms) * 512 * 16 * 80) != cudaSuccess) { printf("ring alloc failed\n"); return 1; }
  run<8, 0>(sms, tiles, d_cycles, d_sink, d_ring, "8 warps, x32 double-buffered");
  run<16, 1>(sms, tiles, d_cycles, d_sink, d_ring, "16 warps, x32 double-buffered");
  run<16, 2>(sms, tiles, d_cycles, d_sink, d_ring, "16 warps, both x32 up front");
  run<16, 3>(sms, tiles, d_cycles, d_sink, d_ring, "16 warps, single x32 buffer");
  run<16, 6>(sms, tiles, d_cycles, d_sink, d_ring, "16 warps, four x16 up front");
  return 0;
}
