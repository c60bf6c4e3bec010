// Micro-probe (not product code): issue-to-completion cost of tcgen05.mma on one SM for the shapes conv_tc uses,
// with the A operand in shared memory (SS) or in TMEM (TS), of tcgen05.cp (shared -> TMEM), and of the two
// interleaved; optionally with four warps hammering shared memory through the LSU at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I led-net_b200/csrc -I include -o /tmp/probe_umma tools/probes/probe_umma.cu
//   gpurun -- '/tmp/... > gpurun_out/probe_umma.txt'
// Output: clk per instruction, from one CTA per SM on all SMs (max over CTAs) and from a single CTA.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace ledb::tc;

namespace ledb { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } }

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// MODE: 0 SS mma, 1 TS mma, 2 cp only, 3 cp + TS mma interleaved (one cp per mma), 4 cp + SS mma interleaved
// NACC: number of accumulators the MMAs rotate over (1 = one dependent chain, like a conv tile's K loop)
template <int MODE, int NACC>
__global__ void __launch_bounds__(192, 1) probe(int n, int reps, int lsu, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) {
    // finite pseudo-random bf16 pairs in [-2, 2) (zeros = data-dependent power would hide a clock drop); lsu == 3: zeros
    uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    const uint32_t lo = 0x3F80u | (x & 0x807Fu), hi16 = 0x3F80u | ((x >> 16) & 0x807Fu);
    reinterpret_cast<uint32_t*>(smem)[i] = (lsu == 3) ? 0u : (lo | (hi16 << 16));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_m128(n);
      const uint32_t hi = desc_hi(8 * 128, 2u);                 // 128 B swizzle, 8-row groups 1024 B apart
      const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 48 * 1024);
      const uint32_t a_tm = tm + 448;                           // A region: columns 448..511
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {                             // 4 K steps x 2 shifted windows
        ad[i] = make_desc(hi, sA + (uint32_t)(i >> 2) * 1280 + (uint32_t)(i & 3) * 32);
        bd[i] = make_desc(hi, sB + (uint32_t)(i & 3) * 32);
      }
      const uint32_t dstep = (NACC > 1) ? (uint32_t)n : 0u;     // NACC * n <= 448 columns
      if (MODE == 8) {
        // conv-like footprint: A = 4 halo slabs of 23 KB (taps at (kh*10+kw)*128 B, SBO 1280), B = 9 resident tap tiles of 8 KB
        const uint32_t hiA = desc_hi(1280, 2u);
        const uint32_t sBt = smem_u32(smem + 96 * 1024);
        unsigned long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        long long t0 = clock64();
        for (int tile = 0; tile < reps / 36; ++tile) {
          if (lsu == 4) { tc_fence_after(); tc_fence_after(); }                 // what the conv MMA thread does per tile
          const uint32_t slab = sA + (uint32_t)(tile & 3) * 23552;
          const uint32_t d = tm + (uint32_t)(tile & 3) * (uint32_t)n * (NACC > 1 ? 1u : 0u);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma(d, make_desc(hiA, slab + (uint32_t)((t / 3) * 10 + (t % 3)) * 128 + k * 32),
                     make_desc(hi, sBt + (uint32_t)t * 8192 + k * 32), idesc, (t | k) ? 1u : 0u);
          }
          tc_commit(&bar2);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        out[blockIdx.x] = (t1 - t0) * 8192 / ((reps / 36) * 36) * reps / 8192;
        out[gridDim.x + blockIdx.x] = (long long)(g1 - g0);
        stop = 1;
        goto done;
      }
      long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t d = tm + (uint32_t)(i % NACC) * dstep;
          if (MODE == 0 || MODE >= 5) tc_mma(d, ad[i], bd[i], idesc, (MODE == 7 && i == 0 && (r % 40) == 0) ? 0u : 1u);
          else if (MODE == 1) mma_ts(d, a_tm + (i & 3) * 8, bd[i], idesc, 1u);
          else if (MODE == 2) cp_128x256b(a_tm + i * 8, ad[i]);
          else if (MODE == 3) { cp_128x256b(a_tm + i * 8, ad[i]); mma_ts(d, a_tm + ((i + 4) & 7) * 8, bd[i], idesc, 1u); }
          else { cp_128x256b(a_tm + i * 8, ad[i]); tc_mma(d, ad[i], bd[i], idesc, 1u); }
        }
        // MODE 5: one tcgen05.commit (to a barrier nobody waits on) per 40 MMAs, like one conv tile; MODE 6: per 8 MMAs;
        // MODE 7: per 40 MMAs + the first MMA of each group overwrites the accumulator (accumulate = 0)
        if (MODE == 6 || ((MODE == 5 || MODE == 7) && (r % 40) == 32)) tc_commit(&bar2);
      }
      tc_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  done:
    __syncwarp();
  } else if (lsu == 2 && warp >= 2) {
    // TMEM read traffic: every warp drains 32 columns of its own lane quadrant (columns the MMAs do not touch),
    // like the conv epilogue does while the next tile's MMAs run
    uint32_t acc = 0;
    long long iters = 0;
    const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    while (!stop) {
      uint32_t v[32];
      tc_ld16(taddr, v);
      tc_ld16(taddr + 16, v + 16);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i];
      ++iters;
    }
    if (acc == 0x12345) out[0] = 0;
    if (threadIdx.x == 64) out[gridDim.x + blockIdx.x] = iters;     // LDTM round trips of warp 2 during the run
  } else if (lsu == 1 && warp >= 2) {
    // LSU traffic: 128-bit shared loads, conflict-free, until the issuing thread is done
    uint32_t acc = 0;
    const uint32_t base = smem_u32(smem) + (uint32_t)(threadIdx.x & 127) * 16;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t a, b, c, d;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + i * 2048));
        acc += a + b + c + d;
      }
    }
    if (acc == 0x12345) out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE, int NACC>
void run(const char* name, int n, int lsu, int sms, long long* d) {
  const int reps = (MODE == 8) ? 36 * 20000 : 8192;      // MODE 8 runs ~20 ms so the clocks settle
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(probe<MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(d, 0, sizeof(long long) * 2 * sms);
  probe<MODE, NACC><<<sms, 192, smem>>>(n, reps, lsu, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, n, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(2 * sms);
  cudaMemcpy(h.data(), d, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("%-22s %5d %5d %4d %12.1f %12.1f", name, n, NACC, lsu, (double)mx / reps, (double)mn / reps);
  if (MODE == 8) printf("   %.0f MHz effective (clock64 / globaltimer), %.1f ns per MMA", 1e3 * (double)h[0] / (double)h[sms], (double)h[sms] / reps);
  if (MODE != 8 && lsu == 2 && h[sms] > 0) printf("   clk per (2 x LDTM.x16 + wait) round trip of one warp: %.0f", (double)h[0] / (double)h[sms]);
  printf("\n");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 2 * sms);
  printf("all %d SMs, one CTA each; clk per loop iteration (max / min over CTAs)\n", sms);
  printf("%-22s %5s %5s %4s %12s %12s\n", "mode", "N", "nacc", "lsu", "clk/iter max", "clk/iter min");
  printf("lsu: 0 = idle, 1 = four warps of ld.shared.v4, 2 = four warps of tcgen05.ld 32x32b.x16 x 2\n");
  for (int lsu = 0; lsu < 3; ++lsu) {
    for (int n : {32, 64, 128, 256}) run<0, 1>("SS mma", n, lsu, sms, d);
    for (int n : {32, 64}) run<0, 4>("SS mma", n, lsu, sms, d);
    if (lsu == 0) {
      for (int n : {32, 64, 128}) run<8, 4>("SS conv-like footprint", n, lsu, sms, d);
      for (int n : {32, 64, 128}) run<8, 4>("  same, zero operands", n, 3, sms, d);
      for (int n : {32, 64, 128}) run<8, 4>("  same, 2 fences / tile", n, 4, sms, d);
      for (int n : {32, 64, 128}) run<5, 1>("SS mma, commit/40", n, lsu, sms, d);
      for (int n : {32, 64, 128}) run<6, 1>("SS mma, commit/8", n, lsu, sms, d);
      for (int n : {32, 64, 128}) run<7, 4>("SS, commit/40, acc=0", n, lsu, sms, d);
    }
    for (int n : {32, 64, 128, 256}) run<1, 1>("TS mma (A in TMEM)", n, lsu, sms, d);
    for (int n : {32, 64}) run<1, 4>("TS mma (A in TMEM)", n, lsu, sms, d);
    run<2, 1>("cp 128x256b", 64, lsu, sms, d);
    for (int n : {32, 64, 128}) run<3, 1>("cp + TS mma", n, lsu, sms, d);
    for (int n : {32, 64, 128}) run<4, 1>("cp + SS mma", n, lsu, sms, d);
  }
  return 0;
}
