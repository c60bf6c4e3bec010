// CUDA-core direct convolution (north_star kernel 2: low-channel / fp32-parity path).
//
// Replaces, for one layer, mmcv ConvModule = Conv2d (+BatchNorm folded) (+ReLU) as used by
//   mmseg/models/utils/basic_block.py:43-57, mmseg/models/backbones/ddrnet.py:68-138,
//   mmseg/models/utils/ppm.py:57-117, mmseg/models/decode_heads/led_head.py:84-99,
// including the pre-activation order ('norm','act','conv'): BN+ReLU applied to the input
// BEFORE zero padding (halo elements are literal zeros), and the residual add + output ReLU of
// BasicBlock/Bottleneck.forward (basic_block.py:62-75, 206-221).
//
// Layout: NHWC activations (generic element strides so the stem can read the caller's NCHW
// tensor or raw uint8 directly), weights [tap][Cin][CoutPad16] fp32, fp32 accumulate.
// Block = 128 threads -> 16x16 output pixels x 16 output channels; each thread owns 2 pixels
// x 16 channels (32 fp32 accumulators).  Cin is staged through shared memory 8 channels at a
// time: input halo tile [8][IH][IW] (conflict-free for stride 1) and weights [tap][8][16].
#include "kernels.h"

namespace ledb {

namespace {

constexpr int TILE = 16;   // output tile edge
constexpr int CK = 8;      // input channels per smem stage
constexpr int COT = 16;    // output channels per block

template <typename Tin>
__device__ __forceinline__ float ld_in(const Tin* p) { return to_f32(__ldg(p)); }

template <typename Tin, typename Tout, int KS>
__global__ void __launch_bounds__(128)
conv_direct_kernel(ConvArgs a) {
  extern __shared__ float smem[];
  const int S = a.stride;
  const int IT = (TILE - 1) * S + (KS - 1) * a.dil + 1;      // input tile edge
  float* s_in = smem;                                        // [CK][IT][IT+pad]
  const int ITP = IT | 1;                                    // odd pitch: fewer bank conflicts for S=2
  float* s_w = smem + CK * IT * ITP;                         // [KS*KS][CK][COT]

  const int tiles_x = (a.Wo + TILE - 1) / TILE;
  const int tx0 = (blockIdx.x % tiles_x) * TILE;
  const int ty0 = (blockIdx.x / tiles_x) * TILE;
  const int co0 = blockIdx.y * COT;
  const int n = blockIdx.z;
  const int t = threadIdx.x;
  const int px = t % TILE, py = t / TILE;                    // pixels (py, px) and (py+8, px)

  const Tin* in = reinterpret_cast<const Tin*>(a.in) + (int64_t)n * a.in_sn;
  const float* wg = a.w_direct;                              // [KS*KS][Cin][CoutPad]
  const int cout_pad = a.cout_pad16;

  float acc[2][COT];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < COT; ++j) acc[i][j] = 0.f;

  const int iy0 = ty0 * S - a.pad, ix0 = tx0 * S - a.pad;

  for (int c0 = 0; c0 < a.Cin; c0 += CK) {
    // ---- stage input halo tile (prologue affine+ReLU applied to in-bounds elements only)
    for (int idx = t; idx < IT * IT; idx += 128) {
      const int y = idx / IT, x = idx % IT;
      const int gy = iy0 + y, gx = ix0 + x;
      bool ok = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W);
      int ry = gy, rx = gx;
      if (a.in_up > 1) {   // virtual zero insertion (dgrad of a strided conv): only multiples of in_up are real
        ok = ok && (gy % a.in_up == 0) && (gx % a.in_up == 0);
        ry = gy / a.in_up; rx = gx / a.in_up;
        ok = ok && ry < a.Hr && rx < a.Wr;
      }
      const Tin* p = in + (int64_t)ry * a.in_sh + (int64_t)rx * a.in_sw;
#pragma unroll
      for (int c = 0; c < CK; ++c) {
        float v = 0.f;
        const int cc = c0 + c;
        if (ok && cc < a.Cin) {
          v = ld_in(p + (int64_t)cc * a.in_sc);
          if (a.pre_scale) {
            v = fmaf(v, a.pre_scale[cc], a.pre_shift[cc]);
            if (a.pre_relu) v = fmaxf(v, 0.f);
          }
        }
        s_in[(c * IT + y) * ITP + x] = v;
      }
    }
    // ---- stage weights [tap][CK][COT]
    for (int idx = t; idx < KS * KS * CK * COT; idx += 128) {
      const int co = idx % COT, c = (idx / COT) % CK, tap = idx / (COT * CK);
      const int cc = c0 + c;
      s_w[idx] = (cc < a.Cin) ? __ldg(wg + ((int64_t)tap * a.Cin + cc) * cout_pad + co0 + co) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kh = 0; kh < KS; ++kh)
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        const float* wt = s_w + (kh * KS + kw) * CK * COT;
        const int sy0 = py * S + kh * a.dil, sy1 = (py + 8) * S + kh * a.dil;
        const int sx = px * S + kw * a.dil;
#pragma unroll
        for (int c = 0; c < CK; ++c) {
          const float a0 = s_in[(c * IT + sy0) * ITP + sx];
          const float a1 = s_in[(c * IT + sy1) * ITP + sx];
          const float4* w4 = reinterpret_cast<const float4*>(wt + c * COT);
#pragma unroll
          for (int q = 0; q < COT / 4; ++q) {
            const float4 w = w4[q];
            acc[0][4 * q + 0] = fmaf(a0, w.x, acc[0][4 * q + 0]);
            acc[0][4 * q + 1] = fmaf(a0, w.y, acc[0][4 * q + 1]);
            acc[0][4 * q + 2] = fmaf(a0, w.z, acc[0][4 * q + 2]);
            acc[0][4 * q + 3] = fmaf(a0, w.w, acc[0][4 * q + 3]);
            acc[1][4 * q + 0] = fmaf(a1, w.x, acc[1][4 * q + 0]);
            acc[1][4 * q + 1] = fmaf(a1, w.y, acc[1][4 * q + 1]);
            acc[1][4 * q + 2] = fmaf(a1, w.z, acc[1][4 * q + 2]);
            acc[1][4 * q + 3] = fmaf(a1, w.w, acc[1][4 * q + 3]);
          }
        }
      }
    __syncthreads();
  }

  // ---- epilogue: bias, residual, ReLU, second (affine+ReLU) output
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int oy = ty0 + py + 8 * i, ox = tx0 + px;
    if (oy >= a.Ho || ox >= a.Wo) continue;
    const int64_t pix = ((int64_t)n * a.Ho + oy) * a.Wo + ox;
    Tout* o = reinterpret_cast<Tout*>(a.out) + pix * a.out_ld;
    Tout* o2 = a.out2 ? reinterpret_cast<Tout*>(a.out2) + pix * a.out2_ld : nullptr;
    const Tout* r = a.res ? reinterpret_cast<const Tout*>(a.res) + pix * a.res_ld : nullptr;
    // fast path: a full, 16 B aligned group of 16 channels -> vector loads/stores
    const bool vec_ok = (co0 + COT <= a.Cout) && (a.out_ld % 8 == 0) && (!o2 || a.out2_ld % 8 == 0) &&
                        (!r || a.res_ld % 8 == 0);
    if (vec_ok) {
      float v[COT], rr[COT];
      if (r) { load8(r + co0, rr); load8(r + co0 + 8, rr + 8); }
#pragma unroll
      for (int j = 0; j < COT; ++j) {
        float t = acc[i][j] + (a.bias ? a.bias[co0 + j] : 0.f);
        if (r) t += rr[j];
        v[j] = a.relu ? fmaxf(t, 0.f) : t;
        if (a.relu == 2) v[j] = fminf(v[j], 6.f);
      }
      if (a.out) { store8(o + co0, v); store8(o + co0 + 8, v + 8); }
      if (o2) {
        float w[COT];
#pragma unroll
        for (int j = 0; j < COT; ++j) {
          const float t = a.o2_scale ? fmaf(v[j], a.o2_scale[co0 + j], a.o2_shift[co0 + j]) : v[j];
          w[j] = fmaxf(t, 0.f);
        }
        store8(o2 + co0, w); store8(o2 + co0 + 8, w + 8);
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < COT; ++j) {
      const int co = co0 + j;
      if (co >= a.Cout) break;
      float v = acc[i][j] + (a.bias ? a.bias[co] : 0.f);
      if (r) v += to_f32(r[co]);
      float vr = a.relu ? fmaxf(v, 0.f) : v;
      if (a.relu == 2) vr = fminf(vr, 6.f);
      if (a.out) o[co] = from_f32<Tout>(vr);
      if (o2) {
        float v2 = a.o2_scale ? fmaf(vr, a.o2_scale[co], a.o2_shift[co]) : vr;
        o2[co] = from_f32<Tout>(fmaxf(v2, 0.f));
      }
    }
  }
}

template <typename Tin, typename Tout>
int launch_t(const ConvArgs& a, cudaStream_t st) {
  const int S = a.stride;
  const int IT = (TILE - 1) * S + (a.ksize - 1) * a.dil + 1;
  const int ITP = IT | 1;
  const size_t smem = sizeof(float) * (CK * IT * ITP + a.ksize * a.ksize * CK * COT);
  dim3 grid(ceil_div(a.Wo, TILE) * ceil_div(a.Ho, TILE), ceil_div(a.Cout, COT), a.N);
  if (a.ksize == 3) {
    if (smem > 48 * 1024)
      LEDB_CUDA_OK(cudaFuncSetAttribute(conv_direct_kernel<Tin, Tout, 3>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_direct_kernel<Tin, Tout, 3><<<grid, 128, smem, st>>>(a);
  } else if (a.ksize == 1) {
    conv_direct_kernel<Tin, Tout, 1><<<grid, 128, smem, st>>>(a);
  } else {
    return fail(LEDB200_EINVAL, "conv_direct: kernel size must be 1 or 3");
  }
  LEDB_LAUNCH_OK("conv_direct_kernel");
  return LEDB200_OK;
}

}  // namespace

int launch_conv_direct(const ConvArgs& a, cudaStream_t st) {
  if (a.in_dtype == LEDB200_F32 && a.out_dtype == LEDB200_F32) return launch_t<float, float>(a, st);
  if (a.in_dtype == LEDB200_BF16 && a.out_dtype == LEDB200_BF16)
    return launch_t<__nv_bfloat16, __nv_bfloat16>(a, st);
  if (a.in_dtype == LEDB200_F32 && a.out_dtype == LEDB200_BF16) return launch_t<float, __nv_bfloat16>(a, st);
  if (a.in_dtype == LEDB200_U8 && a.out_dtype == LEDB200_F32) return launch_t<uint8_t, float>(a, st);
  if (a.in_dtype == LEDB200_U8 && a.out_dtype == LEDB200_BF16) return launch_t<uint8_t, __nv_bfloat16>(a, st);
  return fail(LEDB200_EINVAL, "conv_direct: unsupported dtype combination");
}

}  // namespace ledb
