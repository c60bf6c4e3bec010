// Stem convolution 0 on the tensor cores: Conv2d(3 -> C, 3x3, stride 2, pad 1) + folded BatchNorm + ReLU,
// reading the caller's image directly (normalised fp32 NCHW, or raw uint8 NCHW / NHWC with the
// SegDataPreProcessor affine as a prologue) and writing NHWC bf16.
//
// Replaces mmcv ConvModule `stem[0]` (mmseg/models/backbones/ddrnet.py:121-131) and, for uint8 input,
// SegDataPreProcessor.forward (mmseg/models/data_preprocessor.py:112-118: BGR->RGB, float, (x-mean)/std);
// its second output relu(s*y+b) is the pre-activation BN+ReLU of LEDHead.head_x1
// (mmseg/models/decode_heads/led_head.py:84-99), stored so that zero padding stays literal zero.
//
// K = 27 is far too small for TMA-fed operands, so the im2col tile is built by the threads: a CTA of
// 128 threads owns 128 output pixels (8 rows x 16 cols); thread m gathers its 27 inputs, converts to
// bf16 and writes row m of the A operand (K padded to 32 = one 64 B row, 64 B swizzle) with four
// conflict-free 16 B shared stores.  One elected thread issues two tcgen05.mma (M=128, N=C, K=16) into
// TMEM; every thread then reads back its own accumulator row, applies bias + ReLU (+ the second
// affine + ReLU), stages bf16 through a warp-private swizzled tile and stores 16 B per lane with
// consecutive lanes on consecutive addresses.  Several CTAs are resident per SM (18 KB shared memory,
// 32 TMEM columns each), which is what hides the gather latency.  HBM-bound by design: 12 B (fp32) or
// 3 B (uint8) in and 2 x 64 B out per output pixel.
#include <mutex>

#include "tc_common.cuh"

namespace ledb {
namespace {

using namespace tc;

__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

constexpr int ST_TW = 16, ST_TH = 8;      // output tile (128 pixels = UMMA M)
constexpr int ST_THREADS = 128;
constexpr int ST_MAXN = 32;

struct StemParams {
  int N, H, W, Ho, Wo, Cout, NP;          // NP = Cout padded to 16 (UMMA N)
  int tiles_w, tiles_h;
  uint32_t total_tiles;
  int64_t in_sn, in_sc, in_sh, in_sw;
  const void* in;
  const __nv_bfloat16* w;                 // [NP][27] bf16, k = (kh*3+kw)*3 + ci
  const float* bias;
  const float* pre_scale; const float* pre_shift;
  __nv_bfloat16* out; int out_ld;
  __nv_bfloat16* out2; int out2_ld;
  const float* o2_scale; const float* o2_shift;
  int relu;
  uint32_t tmem_cols;
};

template <typename Tin>
__global__ void __launch_bounds__(ST_THREADS, 6) stem_tc_kernel(const __grid_constant__ StemParams P) {
  __shared__ __align__(1024) uint8_t sA[128 * 64];          // A operand: 128 rows x 32 bf16, 64 B swizzle
  __shared__ __align__(1024) uint8_t sB[ST_MAXN * 64];      // B operand: NP rows x 32 bf16, 64 B swizzle
  __shared__ __align__(1024) uint8_t sStage[2][4][32 * ST_MAXN * 2];   // [output][warp][32 px x NP bf16]
  __shared__ float s_bias[ST_MAXN], s_o2s[ST_MAXN], s_o2b[ST_MAXN], s_ps[4], s_pb[4];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) { mbar_init(&mma_bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, P.tmem_cols);
  // weights -> swizzled K-major B rows (k >= 27 and n >= Cout are zero)
  for (int i = t; i < P.NP * 32; i += ST_THREADS) {
    const int n = i >> 5, k = i & 31;
    const __nv_bfloat16 v = (k < 27) ? P.w[n * 27 + k] : __float2bfloat16(0.f);
    const uint32_t off = (uint32_t)n * 64 + (uint32_t)(((k >> 3) ^ ((n >> 1) & 3)) << 4) + (uint32_t)(k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = v;
  }
  if (t < ST_MAXN) {
    const bool in = t < P.Cout;
    s_bias[t] = (P.bias && in) ? P.bias[t] : 0.f;
    s_o2s[t] = in ? (P.o2_scale ? P.o2_scale[t] : 1.f) : 0.f;
    s_o2b[t] = (P.o2_shift && in) ? P.o2_shift[t] : 0.f;
  }
  if (t < 3) { s_ps[t] = P.pre_scale ? P.pre_scale[t] : 1.f; s_pb[t] = P.pre_shift ? P.pre_shift[t] : 0.f; }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = make_idesc_bf16_m128(P.NP);
  const uint32_t hi = desc_hi(8 * 64, 4u);                  // SBO = 8 rows x 64 B, 64 B swizzle
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
  const float ps0 = s_ps[0], ps1 = s_ps[1], ps2 = s_ps[2], pb0 = s_pb[0], pb1 = s_pb[1], pb2 = s_pb[2];

  const int pr = t >> 4, pc = t & 15;                       // this thread's pixel inside the tile
  const Tin* in = reinterpret_cast<const Tin*>(P.in);
  const int cout8 = (P.Cout + 7) & ~7;
  const int sh = 31 - __clz(P.NP * 2);                      // log2(bytes per staged pixel); NP is 16 or 32
  uint32_t phase = 0;
  const bool direct = P.NP == 32 && P.Cout == 32 && (P.out_ld | P.out2_ld) % 16 == 0 &&
                      (((uintptr_t)P.out | (uintptr_t)P.out2) % 32 == 0);

  for (uint32_t tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % (uint32_t)P.tiles_w);
    const uint32_t r = tile / (uint32_t)P.tiles_w;
    const int th = (int)(r % (uint32_t)P.tiles_h);
    const int n = (int)(r / (uint32_t)P.tiles_h);
    const int oh = th * ST_TH + pr, ow = tw * ST_TW + pc;
    // ---- gather 27 inputs (zero outside the image: padding is applied after normalisation)
    float x[27];
    {
      const int ih0 = 2 * oh - 1, iw0 = 2 * ow - 1;
      const Tin* base = in + (int64_t)n * P.in_sn + (int64_t)ih0 * P.in_sh + (int64_t)iw0 * P.in_sw;
      const bool interior = ih0 >= 0 && iw0 >= 0 && ih0 + 2 < P.H && iw0 + 2 < P.W;   // implies oh < Ho, ow < Wo
      if (interior) {
        // nine (channel, row) pointers, three taps each: no per-tap bounds or address arithmetic
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float sc = ci == 0 ? ps0 : (ci == 1 ? ps1 : ps2), sb = ci == 0 ? pb0 : (ci == 1 ? pb1 : pb2);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const Tin* rp = base + ci * P.in_sc + kh * P.in_sh;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) x[(kh * 3 + kw) * 3 + ci] = fmaf(to_f32(__ldg(rp + kw * P.in_sw)), sc, sb);
          }
        }
      } else {
        const bool pv = oh < P.Ho && ow < P.Wo;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int ih = ih0 + kh;
          const bool rok = pv && ih >= 0 && ih < P.H;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int iw = iw0 + kw;
            const bool ok = rok && iw >= 0 && iw < P.W;
            const Tin* p = base + (int64_t)kh * P.in_sh + (int64_t)kw * P.in_sw;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (ok) {
              v0 = fmaf(to_f32(__ldg(p)), ps0, pb0);
              v1 = fmaf(to_f32(__ldg(p + P.in_sc)), ps1, pb1);
              v2 = fmaf(to_f32(__ldg(p + 2 * P.in_sc)), ps2, pb2);
            }
            x[(kh * 3 + kw) * 3 + 0] = v0; x[(kh * 3 + kw) * 3 + 1] = v1; x[(kh * 3 + kw) * 3 + 2] = v2;
          }
        }
      }
    }
    // ---- row t of the A operand: 4 chunks of 8 bf16, chunk c at ((c ^ ((t>>1)&3)) << 4)
    {
      const uint32_t rowb = (uint32_t)t * 64, sx = (uint32_t)((t >> 1) & 3);
      *reinterpret_cast<uint4*>(sA + rowb + ((0u ^ sx) << 4)) =
          make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
      *reinterpret_cast<uint4*>(sA + rowb + ((1u ^ sx) << 4)) =
          make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15]));
      *reinterpret_cast<uint4*>(sA + rowb + ((2u ^ sx) << 4)) =
          make_uint4(pack_bf16x2(x[16], x[17]), pack_bf16x2(x[18], x[19]), pack_bf16x2(x[20], x[21]), pack_bf16x2(x[22], x[23]));
      *reinterpret_cast<uint4*>(sA + rowb + ((3u ^ sx) << 4)) =
          make_uint4(pack_bf16x2(x[24], x[25]), pack_bf16x2(x[26], 0.f), 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();          // A complete; every thread finished reading TMEM of the previous tile
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        tc_mma(tmem_base, make_desc(hi, sA_u), make_desc(hi, sB_u), idesc, 0u);
        tc_mma(tmem_base, make_desc(hi, sA_u + 32), make_desc(hi, sB_u + 32), idesc, 1u);
        tc_commit(&mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: thread t owns accumulator row t (TMEM lane t)
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (direct) {
      // 32 channels = 64 contiguous bytes per pixel and output: two 256-bit stores each, no staging
      // (same reasoning as the conv_tc fast path: fewer instructions, no shared-memory round trip)
      uint32_t v[32];
      tc_ld16(taddr, v);
      tc_ld16(taddr + 16, v + 16);
      tc_wait_ld();
      if (oh < P.Ho && ow < P.Wo) {
        const int64_t gp = ((int64_t)n * P.Ho + oh) * P.Wo + ow;
        uint4 o[4], o2[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[j] = __uint_as_float(v[8 * g + j]) + s_bias[8 * g + j];
            if (P.relu) f[j] = fmaxf(f[j], 0.f);
          }
          o[g] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          if (P.out2) {
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = fmaxf(fmaf(f[j], s_o2s[8 * g + j], s_o2b[8 * g + j]), 0.f);
            o2[g] = make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]), pack_bf16x2(w[6], w[7]));
          }
        }
        if (P.out) { stg256(P.out + gp * P.out_ld, o[0], o[1]); stg256(P.out + gp * P.out_ld + 16, o[2], o[3]); }
        if (P.out2) { stg256(P.out2 + gp * P.out2_ld, o2[0], o2[1]); stg256(P.out2 + gp * P.out2_ld + 16, o2[2], o2[3]); }
      }
      continue;     // the __syncthreads() after the next tile's gather orders these TMEM reads before its MMAs
    }
    uint8_t* st1 = sStage[0][warp];
    uint8_t* st2 = sStage[1][warp];
    for (int c0 = 0; c0 < P.NP; c0 += 16) {
      uint32_t v[16];
      tc_ld16(taddr + c0, v);
      tc_wait_ld();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = __uint_as_float(v[8 * g + j]) + s_bias[c0 + 8 * g + j];
          if (P.relu) f[j] = fmaxf(f[j], 0.f);
        }
        const uint32_t so = swz(((uint32_t)lane << sh) + (uint32_t)(c0 + 8 * g) * 2);
        if (P.out)
          *reinterpret_cast<uint4*>(st1 + so) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                           pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        if (P.out2) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(f[j], s_o2s[c0 + 8 * g + j], s_o2b[c0 + 8 * g + j]), 0.f);
          *reinterpret_cast<uint4*>(st2 + so) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                           pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
      }
    }
    __syncwarp();
    {
      // warp w staged tile rows 2w, 2w+1 (16 px each); lane l of store i covers staged bytes [512 i + 16 l, +16)
      const int64_t pix0 = ((int64_t)n * P.Ho + th * ST_TH) * P.Wo + tw * ST_TW;
      const int niter = P.NP >> 3;
      const uint32_t bmask = (1u << sh) - 1;
      for (int i = 0; i < niter; ++i) {
        const uint32_t L = (uint32_t)i * 512 + (uint32_t)lane * 16;
        const int p = (int)(L >> sh);
        const int c = (int)((L & bmask) >> 1);
        const int prow = 2 * warp + (p >> 4), pcol = p & 15;
        if (th * ST_TH + prow < P.Ho && tw * ST_TW + pcol < P.Wo && c < cout8) {
          const uint32_t so = swz(L);
          const int64_t gp = pix0 + (int64_t)prow * P.Wo + pcol;
          if (P.out) *reinterpret_cast<uint4*>(P.out + gp * P.out_ld + c) = *reinterpret_cast<const uint4*>(st1 + so);
          if (P.out2) *reinterpret_cast<uint4*>(P.out2 + gp * P.out2_ld + c) = *reinterpret_cast<const uint4*>(st2 + so);
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, P.tmem_cols); }
}

}  // namespace

bool stem_tc_eligible(const ConvArgs& a) {
  if (a.out_dtype != LEDB200_BF16) return false;
  if (a.in_dtype != LEDB200_F32 && a.in_dtype != LEDB200_U8) return false;
  if (a.Cin != 3 || a.ksize != 3 || a.stride != 2 || a.pad != 1 || a.dil != 1) return false;
  if (a.Cout != 16 && a.Cout != 32) return false;
  if (a.res || !a.w_tc) return false;
  if (a.pre_scale && a.pre_relu) return false;
  if (a.out && a.out_ld % 8) return false;
  if (a.out2 && a.out2_ld % 8) return false;
  if ((int64_t)a.N * ceil_div(a.Ho, ST_TH) * ceil_div(a.Wo, ST_TW) >= (1ll << 31)) return false;
  return true;
}

int launch_stem_tc(const ConvArgs& a, cudaStream_t st) {
  if (!stem_tc_eligible(a)) return fail(LEDB200_EINVAL, "stem_tc: shape not eligible");
  StemParams P{};
  P.N = a.N; P.H = a.H; P.W = a.W; P.Ho = a.Ho; P.Wo = a.Wo; P.Cout = a.Cout; P.NP = a.Cout;
  P.tiles_w = ceil_div(a.Wo, ST_TW); P.tiles_h = ceil_div(a.Ho, ST_TH);
  P.total_tiles = (uint32_t)((int64_t)a.N * P.tiles_w * P.tiles_h);
  P.in_sn = a.in_sn; P.in_sc = a.in_sc; P.in_sh = a.in_sh; P.in_sw = a.in_sw;
  P.in = a.in; P.w = a.w_tc; P.bias = a.bias; P.pre_scale = a.pre_scale; P.pre_shift = a.pre_shift;
  P.out = (__nv_bfloat16*)a.out; P.out_ld = a.out_ld; P.out2 = (__nv_bfloat16*)a.out2; P.out2_ld = a.out2_ld;
  P.o2_scale = a.o2_scale; P.o2_shift = a.o2_shift; P.relu = a.relu;
  P.tmem_cols = 32;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int grid = (int)std::min<int64_t>(P.total_tiles, (int64_t)sms * 6);
  if (a.in_dtype == LEDB200_F32) stem_tc_kernel<float><<<grid, ST_THREADS, 0, st>>>(P);
  else stem_tc_kernel<uint8_t><<<grid, ST_THREADS, 0, st>>>(P);
  LEDB_LAUNCH_OK("stem_tc_kernel");
  return LEDB200_OK;
}

}  // namespace ledb
