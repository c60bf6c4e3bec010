// The whole DAPPM (mmseg/models/utils/ppm.py:57-130) in TWO launches instead of twenty-two.
//
// Round 1 ran the module as 22 kernels of < 1 us of work each (bench_ops r1g: 0.27 ms, 0.3-17 % of roofline): every
// conv_tc launch pays ~8 us of ramp + drain (profiles/r1c_notes.md section 5), and its tensors (<= 12 MB) never leave L2.
//
//   dappm_pool_kernel   scales[1..4] up to their 1x1 conv: AvgPool2d(5,2,2) / (9,4,4) / (17,8,8) / global average
//                       (count_include_pad=True, ppm.py:68-90) as column sums + row sums in shared memory, the
//                       pre-activation BN + ReLU, and the 512 -> 128 conv on the few pooled pixels (CUDA cores: 169
//                       pixels per 16 x 32 image).  One CTA per (image, scale, pooled row).
//   dappm_fused_kernel  everything at full DAPPM resolution, one CTA per 16 x 8 pixel tile, one thread-block CLUSTER per
//                       image (<= 8 tiles); tcgen05 MMAs (M = 128 pixels, N = 128) with fp32 accumulators in TMEM:
//        step 0   f0 = scales[0](x) and shortcut(x): A tiles built by 128 worker threads from ONE read of x (the two
//                 pre-activation BN + ReLU applied in registers, st.shared in the 128 B-swizzled K-major layout);
//        step i   f_i = processes[i-1](up(s_i) + f_{i-1}), i = 1..4: the epilogue of step i-1 forms
//                 t_i = relu(bn(f_{i-1} + bilinear(s_i))) (f stays fp32, never stored) and writes it to a ping-pong
//                 buffer in L2; after a cluster barrier the 3x3 conv reads it back by TMA as halo slabs (zero fill =
//                 conv padding, nine shifted UMMA descriptors per slab, as conv_tc.cu);
//        concat + compression + shortcut add: every f_i goes through compression's BN + ReLU slice into a shared-
//                 memory A tile and is multiplied by its 128-column slice of the compression weights into a SECOND
//                 accumulator, which the shortcut conv also accumulates into: the 640-channel concat never exists.
//   Weights stream through a TMA ring (each CTA needs all 2.1 MB of them once).
// Numerics: bf16 operands, fp32 accumulation, like the rest of the bf16 engine; f_i and the concat slices are NOT
// rounded to bf16 between layers (the 22-launch path did), t_i and the output are.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "tc_common.cuh"

namespace ledb {
namespace {

using namespace tc;

constexpr int DTH = 16, DTW = 8;              // pixel tile = UMMA M
constexpr int DP = 128;                       // ppm channels (N of the chain GEMMs) and output channels
constexpr int D_WORKERS = 256;                // warps 0-7: A-tile builders and epilogue (two per TMEM lane quadrant)
constexpr int D_THREADS = 352;                // + warp 8 (weight TMA) + warp 9 (TMEM owner, MMA issuer) + warp 10 (slab TMA)
constexpr uint32_t D_ABOX = (DTH + 2) * (DTW + 2) * 128;   // 23040: halo slab of 64 channels
constexpr uint32_t D_ASTAGE = 24576;
constexpr uint32_t D_TILE = 16384;            // 128 rows x 128 B (built A tile, weight tile, half a concat slice)
constexpr int D_SA = 2, D_SB = 7;           // 48 KB of A stages, 112 KB of weight tiles in flight
constexpr int D_MAXC = 1024;

struct DappmParams {
  int N, H, W, C;                             // x: [N,H,W,C] bf16, C = 16 * channels
  int tiles_w, tiles;                         // tiles per image = cluster size
  const __nv_bfloat16* x;
  const __nv_bfloat16* s[4];                  // pooled branch outputs [N,sh,sw,128] bf16 (dappm_pool_kernel)
  int sh[4], sw[4];
  __nv_bfloat16* T[2];                        // ping-pong t_i [N,H,W,128] bf16
  __nv_bfloat16* out; int out_ld;             // [N,H,W,128]
  // folded pre-activation BN (y = relu(a x + b)) and conv biases, device fp32 (bias pointers may be null)
  const float *a_s0, *b_s0, *a_sc, *b_sc;     // [C]
  const float *a_p[4], *b_p[4];               // [128] each: processes[i].bn
  const float *a_c, *b_c;                     // [640] compression.bn
  const float *bias_s0, *bias_p[4], *bias_c, *bias_sc;
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
// generic-proxy global writes -> visible to other CTAs' TMA (async proxy) reads after the cluster barrier
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// shared-memory float tables (offsets in floats)
struct Tab {
  int a_s0, b_s0, a_sc, b_sc;                 // [C] each
  int a_p, b_p;                               // [4][128]
  int a_c, b_c;                               // [640]
  int bias_s0, bias_p, bias_out;              // [128], [4][128], [128] (compression + shortcut bias)
  int total;
};
__host__ __device__ inline Tab make_tab(int C) {
  Tab t;
  int o = 0;
  t.a_s0 = o; o += C; t.b_s0 = o; o += C; t.a_sc = o; o += C; t.b_sc = o; o += C;
  t.a_p = o; o += 4 * DP; t.b_p = o; o += 4 * DP;
  t.a_c = o; o += 5 * DP; t.b_c = o; o += 5 * DP;
  t.bias_s0 = o; o += DP; t.bias_p = o; o += 4 * DP; t.bias_out = o; o += DP;
  t.total = o;
  return t;
}

__global__ void __launch_bounds__(D_THREADS, 1)
dappm_fused_kernel(const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1,
                   const __grid_constant__ CUtensorMap tmWs0, const __grid_constant__ CUtensorMap tmWsc,
                   const __grid_constant__ CUtensorMap tmWp0, const __grid_constant__ CUtensorMap tmWp1,
                   const __grid_constant__ CUtensorMap tmWp2, const __grid_constant__ CUtensorMap tmWp3,
                   const __grid_constant__ CUtensorMap tmWc, const __grid_constant__ DappmParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                        // [D_SA][D_ASTAGE]
  uint8_t* sB = sA + (size_t)D_SA * D_ASTAGE;                // [D_SB][D_TILE]
  uint8_t* sCat = sB + (size_t)D_SB * D_TILE;                // [2][D_TILE]: one 128-channel concat slice, K-major
  float* tab = reinterpret_cast<float*>(sCat + 2 * D_TILE);
  const Tab TB = make_tab(P.C);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + ((TB.total + 3) & ~3));
  uint64_t* a_full = bars;                  // [D_SA]
  uint64_t* a_empty = a_full + D_SA;        // [D_SA]
  uint64_t* b_full = a_empty + D_SA;        // [D_SB]
  uint64_t* b_empty = b_full + D_SB;        // [D_SB]
  uint64_t* f_full = b_empty + D_SB;        // chain accumulator complete
  uint64_t* cat_full = f_full + 1;          // concat slice written (and the chain accumulator drained)
  uint64_t* cat_empty = cat_full + 1;       // compression MMAs of the slice retired
  uint64_t* c_full = cat_empty + 1;         // output accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = (int)(blockIdx.x % (unsigned)P.tiles), n = (int)(blockIdx.x / (unsigned)P.tiles);
  const int h0 = (tile / P.tiles_w) * DTH, w0 = (tile % P.tiles_w) * DTW;
  const int nkc = P.C / 64;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmT0); prefetch_tensormap(&tmT1); prefetch_tensormap(&tmWs0); prefetch_tensormap(&tmWsc);
    prefetch_tensormap(&tmWp0); prefetch_tensormap(&tmWp1); prefetch_tensormap(&tmWp2); prefetch_tensormap(&tmWp3);
    prefetch_tensormap(&tmWc);
    for (int i = 0; i < D_SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < D_SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(f_full, 1); mbar_init(cat_full, 1); mbar_init(cat_empty, 1); mbar_init(c_full, 1);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  // constant tables
  for (int i = threadIdx.x; i < P.C; i += D_THREADS) {
    tab[TB.a_s0 + i] = P.a_s0[i]; tab[TB.b_s0 + i] = P.b_s0[i];
    tab[TB.a_sc + i] = P.a_sc[i]; tab[TB.b_sc + i] = P.b_sc[i];
  }
  for (int i = threadIdx.x; i < 4 * DP; i += D_THREADS) {
    const int j = i / DP, c = i % DP;
    tab[TB.a_p + i] = P.a_p[j][c]; tab[TB.b_p + i] = P.b_p[j][c];
    tab[TB.bias_p + i] = P.bias_p[j] ? P.bias_p[j][c] : 0.f;
  }
  for (int i = threadIdx.x; i < 5 * DP; i += D_THREADS) { tab[TB.a_c + i] = P.a_c[i]; tab[TB.b_c + i] = P.b_c[i]; }
  for (int i = threadIdx.x; i < DP; i += D_THREADS) {
    tab[TB.bias_s0 + i] = P.bias_s0 ? P.bias_s0[i] : 0.f;
    tab[TB.bias_out + i] = (P.bias_c ? P.bias_c[i] : 0.f) + (P.bias_sc ? P.bias_sc[i] : 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_f = tmem_base, acc_c = tmem_base + DP;

  // ring positions: every role walks the same item sequence
  int ia = 0, ib = 0;                                    // items consumed / produced so far
  auto a_stage = [&](int it) { return it % D_SA; };
  auto a_par = [&](int it) { return (it / D_SA) & 1; };
  auto b_stage = [&](int it) { return it % D_SB; };
  auto b_par = [&](int it) { return (it / D_SB) & 1; };

  if (warp < 8) {
    // =========================== workers (8 warps): A tiles of step 0, every epilogue =================
    // Two warps per TMEM lane quadrant: thread (m, half) owns tile pixel m and channel half `half` of every 64- or
    // 128-channel row.  (Four warps - one per scheduler, nothing to hide a dependent-issue latency behind - made the
    // epilogues the critical path of the kernel: ncu r2k, 88 us at 6.6 clk per warp instruction.)
    const int m = threadIdx.x & 127, half = threadIdx.x >> 7, ph = m >> 3, pw = m & 7;
    const int y = h0 + ph, x = w0 + pw;
    const bool valid = y < P.H && x < P.W;
    const int64_t pix = ((int64_t)n * P.H + y) * P.W + x;
    const uint32_t row_off = (uint32_t)m * 128, sw = (uint32_t)(m & 7);
    const uint32_t sA_u = smem_u32(sA), sCat_u = smem_u32(sCat);
    const uint32_t taddr_q = ((uint32_t)((warp & 3) * 32) << 16);
    auto affine_relu8 = [](float* f, const float* sa, const float* sb) {
      const float4 a0 = *reinterpret_cast<const float4*>(sa), a1 = *reinterpret_cast<const float4*>(sa + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sb), b1 = *reinterpret_cast<const float4*>(sb + 4);
      f[0] = fmaxf(fmaf(f[0], a0.x, b0.x), 0.f); f[1] = fmaxf(fmaf(f[1], a0.y, b0.y), 0.f);
      f[2] = fmaxf(fmaf(f[2], a0.z, b0.z), 0.f); f[3] = fmaxf(fmaf(f[3], a0.w, b0.w), 0.f);
      f[4] = fmaxf(fmaf(f[4], a1.x, b1.x), 0.f); f[5] = fmaxf(fmaf(f[5], a1.y, b1.y), 0.f);
      f[6] = fmaxf(fmaf(f[6], a1.z, b1.z), 0.f); f[7] = fmaxf(fmaf(f[7], a1.w, b1.w), 0.f);
    };
    // ---- step 0: one read of x feeds scales[0] (BN s0) and shortcut (BN sc)
    {
      const uint4* xp = reinterpret_cast<const uint4*>(P.x + pix * P.C) + half * 4;
      uint4 cur[4], nxt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = valid ? __ldg(xp + j) : make_uint4(0, 0, 0, 0);
      for (int kc = 0; kc < nkc; ++kc) {
        if (kc + 1 < nkc) {
#pragma unroll
          for (int j = 0; j < 4; ++j) nxt[j] = valid ? __ldg(xp + (kc + 1) * 8 + j) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const float* sa = tab + (which ? TB.a_sc : TB.a_s0) + kc * 64 + half * 32;
          const float* sb = tab + (which ? TB.b_sc : TB.b_s0) + kc * 64 + half * 32;
          const int st = a_stage(ia);
          mbar_wait(&a_empty[st], a_par(ia) ^ 1);
          const uint32_t dst = sA_u + (uint32_t)st * D_ASTAGE + row_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
            unpack8(cur[j], f);
            affine_relu8(f, sa + 8 * j, sb + 8 * j);
            // pixels outside the image contribute nothing (their outputs are never stored)
            sts128(dst + (((uint32_t)(half * 4 + j) ^ sw) << 4), valid ? pack8(f) : make_uint4(0, 0, 0, 0));
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (threadIdx.x == 0) mbar_arrive(&a_full[st]);
          ++ia;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
      }
    }
    // ---- epilogues of steps 0..4
    for (int s = 0; s < 5; ++s) {
      // up(s_{s+1}) at this pixel for this thread's 64 channels, formed while the step's MMAs run (ATen
      // align_corners=False, the arithmetic of upsample_add_kernel)
      float u[64];
      if (s < 4) {
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        const int sh_ = P.sh[s], sw_ = P.sw[s];
        bilinear_coord(min(y, P.H - 1), (float)sh_ / (float)P.H, sh_, y0, y1, ly0, ly1);
        bilinear_coord(min(x, P.W - 1), (float)sw_ / (float)P.W, sw_, x0, x1, lx0, lx1);
        const __nv_bfloat16* sp = P.s[s] + (int64_t)n * sh_ * sw_ * DP + half * 64;
        const __nv_bfloat16 *p00 = sp + (int64_t)(y0 * sw_ + x0) * DP, *p01 = sp + (int64_t)(y0 * sw_ + x1) * DP;
        const __nv_bfloat16 *p10 = sp + (int64_t)(y1 * sw_ + x0) * DP, *p11 = sp + (int64_t)(y1 * sw_ + x1) * DP;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float a[8], b[8], cc[8], d[8];
          load8(p00 + 8 * g, a); load8(p01 + 8 * g, b); load8(p10 + 8 * g, cc); load8(p11 + 8 * g, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float r0 = fmaf(b[j], lx1, a[j] * lx0);
            const float r1 = fmaf(d[j], lx1, cc[j] * lx0);
            u[8 * g + j] = fmaf(r1, ly1, r0 * ly0);
          }
        }
      }
      const float* bias = tab + (s == 0 ? TB.bias_s0 : TB.bias_p + (s - 1) * DP) + half * 64;
      const float* ac = tab + TB.a_c + s * DP + half * 64;
      const float* bc = tab + TB.b_c + s * DP + half * 64;
      const float* ap = tab + TB.a_p + s * DP + half * 64;       // processes[s].bn (s < 4)
      const float* bp = tab + TB.b_p + s * DP + half * 64;
      __nv_bfloat16* tdst = P.T[s & 1] + pix * DP + half * 64;
      mbar_wait(f_full, s & 1);
      tc_fence_after();
      if (s > 0) mbar_wait(cat_empty, (s - 1) & 1);    // the previous slice's compression MMAs have retired
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t v[32];
        tc_ld16(acc_f + taddr_q + half * 64 + cb * 32, v);
        tc_ld16(acc_f + taddr_q + half * 64 + cb * 32 + 16, v + 16);
        tc_wait_ld();
        uint4 to[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = cb * 32 + 8 * g;               // channel inside this thread's half
          float f[8], o[8];
          const float4 q0 = *reinterpret_cast<const float4*>(bias + c), q1 = *reinterpret_cast<const float4*>(bias + c + 4);
          f[0] = __uint_as_float(v[8 * g + 0]) + q0.x; f[1] = __uint_as_float(v[8 * g + 1]) + q0.y;
          f[2] = __uint_as_float(v[8 * g + 2]) + q0.z; f[3] = __uint_as_float(v[8 * g + 3]) + q0.w;
          f[4] = __uint_as_float(v[8 * g + 4]) + q1.x; f[5] = __uint_as_float(v[8 * g + 5]) + q1.y;
          f[6] = __uint_as_float(v[8 * g + 6]) + q1.z; f[7] = __uint_as_float(v[8 * g + 7]) + q1.w;
          // concat slice through compression's BN + ReLU -> A tile of the compression GEMM (k-chunk = channel half)
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = f[j];
          affine_relu8(o, ac + c, bc + c);
          sts128(sCat_u + (uint32_t)half * D_TILE + row_off + ((((uint32_t)c >> 3) ^ sw) << 4), pack8(o));
          if (s < 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = f[j] + u[c + j];
            affine_relu8(o, ap + c, bp + c);
            to[g] = pack8(o);
          }
        }
        if (s < 4 && valid) { stg256(tdst + cb * 32, to[0], to[1]); stg256(tdst + cb * 32 + 16, to[2], to[3]); }
      }
      tc_fence_before();
      fence_proxy_async_all();          // concat slice (shared) -> UMMA, t_{s+1} (global) -> the cluster's TMA reads
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) mbar_arrive(cat_full);
      if (s < 4) { cluster_arrive(); cluster_wait(); }
    }
    // ---- output: compression(concat) + shortcut(x) (+ both biases)
    {
      mbar_wait(c_full, 0);
      tc_fence_after();
      const float* bo = tab + TB.bias_out + half * 64;
      __nv_bfloat16* op = P.out + pix * P.out_ld + half * 64;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t v[32];
        tc_ld16(acc_c + taddr_q + half * 64 + cb * 32, v);
        tc_ld16(acc_c + taddr_q + half * 64 + cb * 32 + 16, v + 16);
        tc_wait_ld();
        uint4 o[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
          const float4 q0 = *reinterpret_cast<const float4*>(bo + cb * 32 + 8 * g), q1 = *reinterpret_cast<const float4*>(bo + cb * 32 + 8 * g + 4);
          f[0] = __uint_as_float(v[8 * g + 0]) + q0.x; f[1] = __uint_as_float(v[8 * g + 1]) + q0.y;
          f[2] = __uint_as_float(v[8 * g + 2]) + q0.z; f[3] = __uint_as_float(v[8 * g + 3]) + q0.w;
          f[4] = __uint_as_float(v[8 * g + 4]) + q1.x; f[5] = __uint_as_float(v[8 * g + 5]) + q1.y;
          f[6] = __uint_as_float(v[8 * g + 6]) + q1.z; f[7] = __uint_as_float(v[8 * g + 7]) + q1.w;
          o[g] = pack8(f);
        }
        if (valid) { stg256(op + cb * 32, o[0], o[1]); stg256(op + cb * 32 + 16, o[2], o[3]); }
      }
    }
  } else if (warp == 8) {
    // =========================== weight producer: runs AHEAD of the cluster barriers ===================
    // Weights do not depend on the other CTAs, so the tiles of step s+1 are requested while step s's epilogue and the
    // barrier run (112 KB in flight).  The warp arrives at barrier s before it starts on step s+1's tiles and only
    // waits for it afterwards: blocking on a full ring in between cannot stall the barrier (arrivals complete it).
    const CUtensorMap* tmWp[4] = {&tmWp0, &tmWp1, &tmWp2, &tmWp3};
    for (int s = 0; s < 5; ++s) {
      if (s > 0) cluster_arrive();
      if (elect_one()) {
        auto load_b = [&](const CUtensorMap* tm, int k0) {
          const int st = b_stage(ib);
          mbar_wait(&b_empty[st], b_par(ib) ^ 1);
          mbar_expect_tx(&b_full[st], D_TILE);
          tma_load_2d(smem_u32(sB + (size_t)st * D_TILE), tm, smem_u32(&b_full[st]), k0, 0);
          ++ib;
        };
        if (s == 0) {
          for (int kc = 0; kc < nkc; ++kc) { load_b(&tmWs0, kc * 64); load_b(&tmWsc, kc * 64); }
        } else {
          for (int ch = 0; ch < 2; ++ch)
            for (int t = 0; t < 9; ++t) load_b(tmWp[s - 1], t * DP + ch * 64);
        }
        load_b(&tmWc, s * DP);
        load_b(&tmWc, s * DP + 64);
      }
      __syncwarp();
      if (s > 0) cluster_wait();
    }
  } else if (warp == 10) {
    // =========================== slab producer: t_s of the whole image, after barrier s-1 =============
    ia = 2 * nkc;                                       // the workers filled the first 2 * nkc A items
    for (int s = 1; s < 5; ++s) {
      cluster_arrive(); cluster_wait();
      if (elect_one()) {
        const CUtensorMap* tmT = ((s - 1) & 1) ? &tmT1 : &tmT0;
        fence_proxy_async_all();                         // reader side of the t_s hand-over (the writers fenced before arriving)
        for (int ch = 0; ch < 2; ++ch) {
          const int st = a_stage(ia);
          mbar_wait(&a_empty[st], a_par(ia) ^ 1);
          mbar_expect_tx(&a_full[st], D_ABOX);
          tma_load_4d(smem_u32(sA + (size_t)st * D_ASTAGE), tmT, smem_u32(&a_full[st]), ch * 64, w0 - 1, h0 - 1, n);
          ++ia;
        }
      }
      __syncwarp();
    }
  } else {
    // =========================== MMA issuer (warp 9) ================================================
  if (warp == 9) {
    const uint32_t idesc = make_idesc_bf16_m128(DP);
    const uint32_t hi_tile = desc_hi(1024, 2u);                       // 8-row groups 1024 B apart, 128 B swizzle
    const uint32_t hi_slab = desc_hi((DTW + 2) * 128, 2u);            // halo slab: one slab row between 8-row groups
    const uint32_t sA16 = (smem_u32(sA) >> 4) & 0x3FFFu, sB16 = (smem_u32(sB) >> 4) & 0x3FFFu;
    const uint32_t sCat16 = (smem_u32(sCat) >> 4) & 0x3FFFu;
    constexpr uint32_t LBO1 = 1u << 16;
    for (int s = 0; s < 5; ++s) {
      if (elect_one()) {
        auto mma4 = [&](uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, bool first_acc0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k == 0 && first_acc0) tc_mma2(d, a_lo, a_hi, b_lo, hi_tile, idesc, 0u);
            else tc_mma2_acc(d, a_lo + 2 * k, a_hi, b_lo + 2 * k, hi_tile, idesc);
          }
        };
        auto wait_b = [&]() -> uint32_t {
          const int st = b_stage(ib);
          mbar_wait(&b_full[st], b_par(ib));
          tc_fence_after();
          return (sB16 + (uint32_t)st * (D_TILE >> 4)) | LBO1;
        };
        auto free_b = [&]() { tc_commit(&b_empty[b_stage(ib)]); ++ib; };
        if (s == 0) {
          for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              const int st = a_stage(ia);
              mbar_wait(&a_full[st], a_par(ia));
              tc_fence_after();
              const uint32_t a_lo = (sA16 + (uint32_t)st * (D_ASTAGE >> 4)) | LBO1;
              const uint32_t b_lo = wait_b();
              mma4(which ? acc_c : acc_f, a_lo, hi_tile, b_lo, kc == 0);
              free_b();
              tc_commit(&a_empty[st]);
              ++ia;
            }
          }
        } else {
          for (int ch = 0; ch < 2; ++ch) {
            const int st = a_stage(ia);
            mbar_wait(&a_full[st], a_par(ia));
            tc_fence_after();
            const uint32_t a_base = (sA16 + (uint32_t)st * (D_ASTAGE >> 4)) | LBO1;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t b_lo = wait_b();
              const uint32_t a_lo = a_base + (uint32_t)(((t / 3) * (DTW + 2) + (t % 3)) * 128 >> 4);
              mma4(acc_f, a_lo, hi_slab, b_lo, ch == 0 && t == 0);
              free_b();
            }
            tc_commit(&a_empty[st]);
            ++ia;
          }
        }
        tc_commit(f_full);
        // compression slice s: A = the concat slice the workers are writing now
        mbar_wait(cat_full, s & 1);
        tc_fence_after();
#pragma unroll
        for (int kc2 = 0; kc2 < 2; ++kc2) {
          const uint32_t b_lo = wait_b();
          const uint32_t a_lo = (sCat16 + (uint32_t)kc2 * (D_TILE >> 4)) | LBO1;
          mma4(acc_c, a_lo, hi_tile, b_lo, false);
          free_b();
        }
        tc_commit(cat_empty);
        if (s == 4) tc_commit(c_full);
      }
      __syncwarp();
      if (s < 4) { cluster_arrive(); cluster_wait(); }
    }
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ pooled branches
struct PoolParams {
  int N, H, W, C;
  const __nv_bfloat16* x;
  int k[4], st[4], pd[4], ph[4], pw[4], row0[4];   // pooling window / stride / pad, pooled size, first block row of the scale
  int rows_total;
  const float *a[4], *b[4];                        // scales[i].1 BN as y = relu(a x + b), [C]
  const __nv_bfloat16* w[4];                       // [128][C] bf16 K-major
  const float* bias[4];
  __nv_bfloat16* out[4];                           // [N,ph,pw,128]
};

constexpr int PK_THREADS = 512;

__global__ void __launch_bounds__(PK_THREADS)
dappm_pool_kernel(const __grid_constant__ PoolParams P) {
  extern __shared__ __align__(16) float psm[];
  float* colsum = psm;                               // [W][C]
  float* pooled = psm + (size_t)P.W * P.C;           // [8][C]
  float* partial = pooled + 8 * P.C;                 // [4][8][128]
  const int n = (int)(blockIdx.x / (unsigned)P.rows_total);
  int r = (int)(blockIdx.x % (unsigned)P.rows_total), sc = 0;
  while (sc < 3 && r >= P.row0[sc + 1]) ++sc;
  const int py = r - P.row0[sc];
  const int k = P.k[sc], s = P.st[sc], pd = P.pd[sc], pw = P.pw[sc];
  const int C = P.C, W = P.W, H = P.H;
  int ya, yb;
  float inv_h;
  if (k == 0) { ya = 0; yb = H; inv_h = 1.f; }
  else {
    const int y0 = py * s - pd, hend = min(y0 + k, H + pd);
    inv_h = (float)(hend - y0);                      // count_include_pad=True: the divisor counts the padding (ATen pool_size)
    ya = max(y0, 0); yb = min(y0 + k, H);
  }
  // ---- column sums over the window's rows: thread = 8 channels of one column (16 B loads), up to 9 rows in flight
  //      (the first version read 4 B per load with one accumulator: 136 dependent L2 round trips per thread, 48 us)
  const __nv_bfloat16* xb = P.x + (int64_t)n * H * W * C;
  for (int i = threadIdx.x; i < W * (C / 8); i += PK_THREADS) {
    const int c8 = i % (C / 8), x = i / (C / 8);
    const uint4* col = reinterpret_cast<const uint4*>(xb + (int64_t)x * C) + c8;
    const int64_t rstride = (int64_t)W * C / 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int y = ya; y < yb; y += 9) {
      uint4 v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = (y + j < yb) ? __ldg(col + (y + j) * rstride) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float f[8];
        unpack8(v[j], f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += f[q];
      }
    }
    float4* dst = reinterpret_cast<float4*>(colsum + x * C + 8 * c8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  const float* a = P.a[sc];
  const float* b = P.b[sc];
  const __nv_bfloat16* wt = P.w[sc];
  const int co = threadIdx.x & 127, half = threadIdx.x >> 7;           // output channel, quarter of the input channels
  for (int px0 = 0; px0 < pw; px0 += 8) {
    const int npx = min(8, pw - px0);
    // ---- pooled pixels of this group: row sums of the column sums, average, BN + ReLU (rounded to bf16 like every
    //      activation the engine stores)
    for (int i = threadIdx.x; i < npx * C; i += PK_THREADS) {
      const int c = i % C, px = px0 + i / C;
      int xa, xe;
      float inv;
      if (k == 0) { xa = 0; xe = W; inv = 1.f / (float)(H * W); }
      else {
        const int x0 = px * s - pd, wend = min(x0 + k, W + pd);
        inv = 1.f / (inv_h * (float)(wend - x0));
        xa = max(x0, 0); xe = min(x0 + k, W);
      }
      float acc = 0.f;
      for (int x = xa; x < xe; ++x) acc += colsum[x * C + c];
      const float v = fmaxf(fmaf(acc * inv, a[c], b[c]), 0.f);
      pooled[(i / C) * C + c] = __bfloat162float(__float2bfloat16_rn(v));
    }
    __syncthreads();
    // ---- 1x1 conv C -> 128 on up to 8 pixels: thread = (output channel, half of K)
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int cbeg = half * (C / 4);
    const uint4* wrow = reinterpret_cast<const uint4*>(wt + (int64_t)co * C + cbeg);
    for (int c8 = 0; c8 < C / 32; ++c8) {
      float wv[8];
      unpack8(__ldg(wrow + c8), wv);
      const float* pp = pooled + cbeg + c8 * 8;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        if (p < npx) {
          const float4 q0 = *reinterpret_cast<const float4*>(pp + p * C), q1 = *reinterpret_cast<const float4*>(pp + p * C + 4);
          acc[p] = fmaf(wv[0], q0.x, acc[p]); acc[p] = fmaf(wv[1], q0.y, acc[p]);
          acc[p] = fmaf(wv[2], q0.z, acc[p]); acc[p] = fmaf(wv[3], q0.w, acc[p]);
          acc[p] = fmaf(wv[4], q1.x, acc[p]); acc[p] = fmaf(wv[5], q1.y, acc[p]);
          acc[p] = fmaf(wv[6], q1.z, acc[p]); acc[p] = fmaf(wv[7], q1.w, acc[p]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) partial[(half * 8 + p) * 128 + co] = acc[p];
    __syncthreads();
    for (int i = threadIdx.x; i < npx * 128; i += PK_THREADS) {
      const int p = i >> 7, c = i & 127;
      const float v = (partial[p * 128 + c] + partial[(8 + p) * 128 + c]) + (partial[(16 + p) * 128 + c] + partial[(24 + p) * 128 + c]) +
                      (P.bias[sc] ? P.bias[sc][c] : 0.f);
      P.out[sc][(((int64_t)n * P.ph[sc] + py) * pw + px0 + p) * 128 + c] = __float2bfloat16_rn(v);
    }
    __syncthreads();
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}
int encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeFn fn = get_encode();
  if (!fn) return fail(LEDB200_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LEDB200_ECUDA, "dappm: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return LEDB200_OK;
}

}  // namespace

bool dappm_eligible(const DappmArgs& a) {
  if (a.C < 64 || a.C % 64 || a.C > D_MAXC) return false;
  if (a.P != DP || a.Cout != DP || a.out_ld < DP || a.out_ld % 16) return false;
  if (a.N < 1 || a.H < 1 || a.W < 1) return false;
  const int tiles = ceil_div(a.H, DTH) * ceil_div(a.W, DTW);
  if (tiles > 8) return false;                                   // one portable cluster per image
  if ((size_t)(a.W * a.C + 8 * a.C + 4 * 8 * 128) * 4 > 200 * 1024) return false;   // pool kernel's column sums
  return true;
}

int launch_dappm(const DappmArgs& a, cudaStream_t st) {
  if (!dappm_eligible(a)) return fail(LEDB200_EINVAL, "dappm: shape not eligible");
  // ---- pooled branches
  {
    PoolParams P{};
    P.N = a.N; P.H = a.H; P.W = a.W; P.C = a.C; P.x = (const __nv_bfloat16*)a.x;
    int rows = 0;
    for (int i = 0; i < 4; ++i) {
      P.k[i] = a.pool_k[i]; P.st[i] = a.pool_s[i]; P.pd[i] = a.pool_p[i]; P.ph[i] = a.sh[i]; P.pw[i] = a.sw[i];
      P.row0[i] = rows; rows += a.sh[i];
      P.a[i] = a.a_scale[i]; P.b[i] = a.b_scale[i]; P.w[i] = a.w_scale[i]; P.bias[i] = a.bias_scale[i];
      P.out[i] = (__nv_bfloat16*)a.s[i];
    }
    P.rows_total = rows;
    const size_t smem = (size_t)(a.W * a.C + 8 * a.C + 4 * 8 * 128) * 4;
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] { err = cudaFuncSetAttribute(dappm_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (err != cudaSuccess) return fail(LEDB200_ECUDA, std::string("dappm: cudaFuncSetAttribute: ") + cudaGetErrorString(err));
    dappm_pool_kernel<<<a.N * rows, PK_THREADS, smem, st>>>(P);
    LEDB_LAUNCH_OK("dappm_pool_kernel");
  }
  // ---- fused chain
  DappmParams P{};
  P.N = a.N; P.H = a.H; P.W = a.W; P.C = a.C;
  P.tiles_w = ceil_div(a.W, DTW); P.tiles = ceil_div(a.H, DTH) * P.tiles_w;
  P.x = (const __nv_bfloat16*)a.x;
  for (int i = 0; i < 4; ++i) {
    P.s[i] = (const __nv_bfloat16*)a.s[i]; P.sh[i] = a.sh[i]; P.sw[i] = a.sw[i];
    P.a_p[i] = a.a_proc[i]; P.b_p[i] = a.b_proc[i]; P.bias_p[i] = a.bias_proc[i];
  }
  P.T[0] = (__nv_bfloat16*)a.t0; P.T[1] = (__nv_bfloat16*)a.t1;
  P.out = (__nv_bfloat16*)a.out; P.out_ld = a.out_ld;
  P.a_s0 = a.a_s0; P.b_s0 = a.b_s0; P.a_sc = a.a_sc; P.b_sc = a.b_sc; P.a_c = a.a_comp; P.b_c = a.b_comp;
  P.bias_s0 = a.bias_s0; P.bias_c = a.bias_comp; P.bias_sc = a.bias_sc;
  CUtensorMap tmT[2], tmWs0, tmWsc, tmWp[4], tmWc;
  int rc;
  for (int i = 0; i < 2; ++i) {
    const uint64_t dims[4] = {(uint64_t)DP, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t str[3] = {(uint64_t)DP * 2, (uint64_t)a.W * DP * 2, (uint64_t)a.H * a.W * DP * 2};
    const uint32_t box[4] = {64, DTW + 2, DTH + 2, 1};
    if ((rc = encode(&tmT[i], i ? a.t1 : a.t0, 4, dims, str, box))) return rc;
  }
  auto enc_w = [&](CUtensorMap* m, const void* w, int K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)DP};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, DP};
    return encode(m, w, 2, dims, str, box);
  };
  if ((rc = enc_w(&tmWs0, a.w_s0, a.C))) return rc;
  if ((rc = enc_w(&tmWsc, a.w_sc, a.C))) return rc;
  for (int i = 0; i < 4; ++i) if ((rc = enc_w(&tmWp[i], a.w_proc[i], 9 * DP))) return rc;
  if ((rc = enc_w(&tmWc, a.w_comp, 5 * DP))) return rc;

  const Tab TB = make_tab(a.C);
  const size_t smem = 1024 + (size_t)D_SA * D_ASTAGE + (size_t)D_SB * D_TILE + 2 * D_TILE + (size_t)((TB.total + 3) & ~3) * 4 +
                      (2 * D_SA + 2 * D_SB + 4) * 8 + 16;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] { err = cudaFuncSetAttribute(dappm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); });
  if (err != cudaSuccess) return fail(LEDB200_ECUDA, std::string("dappm: cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  if (smem > 226 * 1024) return fail(LEDB200_EINVAL, "dappm: shared-memory plan does not fit");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(a.N * P.tiles), 1, 1);
  cfg.blockDim = dim3(D_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.tiles; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, dappm_fused_kernel, tmT[0], tmT[1], tmWs0, tmWsc, tmWp[0], tmWp[1], tmWp[2], tmWp[3],
                                     tmWc, P);
  if (e != cudaSuccess) return fail(LEDB200_ECUDA, std::string("dappm_fused_kernel: ") + cudaGetErrorString(e));
  return LEDB200_OK;
}

}  // namespace ledb
