// Training-step kernels (north_star kernel 6): the backward of every op
// on the LED-Net path plus the train-mode forward pieces that differ from inference.
//
// The reference trains through autograd over stock ATen/cuDNN ops
//   (mmseg/models/segmentors/encoder_decoder.py:161-185 loss -> decode_head.loss ->
//    led_head.py:101-146; optimiser SGD lr 0.01 m 0.9 wd 5e-4, configs/LED_Net/...py:64-65).
// Here each op is one or two hand-written kernels, called from torch.autograd.Function wrappers
// (led-net_b200/train_ops.py); PyTorch supplies the tape, the memory and NCCL only.
//
//   conv forward / data gradient .. conv_direct.cu through device-side weight packing below; the data
//                                   gradient of a stride-2 conv is a stride-1 conv of the flipped,
//                                   transposed weights over a zero-inserted view of dY.
//   conv weight gradient .......... wgrad_kernel: dW[co,ci,kh,kw] = sum_{n,y,x} dY[n,y,x,co] X[n,y*s+kh-p,x*s+kw-p,ci]
//   BatchNorm (training) .......... batch statistics (double accumulation), normalise (+residual)(+ReLU),
//                                   running-stat update; backward = one reduction (dgamma, dbeta) + one
//                                   elementwise pass (dX, dResidual) with the ReLU mask recomputed from the output.
//   bilinear resize ............... gather forward AND gather backward (each source pixel sums the outputs that read it, in a
//                                   fixed order), align_corners=False (wrappers.py:8-27)
//   Every reduction is order-fixed (no floating-point atomics): a step is bit-reproducible run to run.
//   add(+ReLU), avg-pool, channel copy (concat/slice), SGD+momentum over one flat parameter arena.
// All tensors NHWC fp32, dense (pixel stride = C) unless an `ld` says otherwise.
#include <algorithm>

#include "tc_common.cuh"

namespace ledb {
namespace {

constexpr int kT = 256;
// tf32 storage mode (ledb200_train_set_tf32_rounding): every activation / gradient tensor these kernels write is rounded to
// tf32 (round to nearest, cvt.rna) so that a tensor-core consumer, which TRUNCATES raw fp32 operands to tf32 (measured:
// tests/test_gpu_train_tc.py, -3.7e-4 mean), sees exactly representable values - truncation would shrink every data
// gradient by ~3e-4 per layer, compounding to percents at the stem.  Off: plain fp32 stores (the CUDA-core path).
int g_round = 0;
// tensor-core passes per product (ledb200_train_set_tf32_passes): 3 = error-compensated 3 x TF32 (fp32-grade, default),
// 1 = single tf32 pass (cuDNN's allow_tf32 numerics; callers then also switch the tf32 storage mode on)
int g_passes = 3;
// passes of the WEIGHT-gradient kernel (ledb200_train_set_wgrad_passes).  A weight gradient is a leaf of the backward pass:
// its rounding is not fed back into the ill-conditioned chain, and it is a sum over 1e5..1e6 pixels in which the operands'
// tf32 truncation shows up as a uniform ~7e-4 shrink plus noise that averages out - two decades inside the 1e-2 gate.
// So it runs ONE tf32 pass by default (a third of the tensor work); 3 keeps it fp32-grade like the other two convolutions.
int g_wgrad_passes = 1;
__device__ __forceinline__ float rt(float v, int rnd) {
  if (!rnd) return v;
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
  return __uint_as_float(t);
}
inline int grid1d(int64_t work, int per_sm = 8) {
  int64_t b = ceil_div64(work, kT);
  const int64_t cap = 148 * per_sm;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// ---------------------------------------------------------------------------------------------
// weight packing, device to device.  w: OIHW fp32 [Cout][Cin][k][k].
//   mode 0 (forward):       out[tap][ci][co_pad16]            = w[co][ci][tap]
//   mode 1 (data gradient): out[tap'][co][ci_pad16], tap' = k*k-1-tap (180 degree rotation), roles swapped
__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int taps,
                                   int mode, int pad16) {
  const int64_t total = (int64_t)taps * (mode ? Cout : Cin) * pad16;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int inner = (int)(i % pad16);
    const int mid = (int)((i / pad16) % (mode ? Cout : Cin));
    const int tap = (int)(i / ((int64_t)pad16 * (mode ? Cout : Cin)));
    float v = 0.f;
    if (mode == 0) {
      if (inner < Cout) v = w[((int64_t)inner * Cin + mid) * taps + tap];
    } else {
      if (inner < Cin) v = w[((int64_t)mid * Cin + inner) * taps + (taps - 1 - tap)];
    }
    out[i] = v;
  }
}

// K-major fp32 weight matrices of the tensor-core path (conv_tc.cu, TF32 instantiations): w_hi = w rounded to tf32 (RN)
// followed by w_lo = w - w_hi, each
//   mode 0 (forward):       out[co_pad][tap * Cin + ci]  = w[co][ci][tap]
//   mode 1 (data gradient): out[ci_pad][tap' * Cout + co] = w[co][ci][taps - 1 - tap']   (rotated, roles swapped)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int taps,
                                      int mode, int rows_pad) {
  const int rows = mode ? Cin : Cout, inner = mode ? Cout : Cin;
  const int64_t total = (int64_t)rows_pad * taps * inner;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % inner);
    const int tap = (int)((i / inner) % taps);
    const int r = (int)(i / ((int64_t)inner * taps));
    float v = 0.f;
    if (r < rows) v = mode == 0 ? w[((int64_t)r * Cin + c) * taps + tap] : w[((int64_t)c * Cin + r) * taps + (taps - 1 - tap)];
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    const float hi = __uint_as_float(t);
    out[i] = hi;                 // rows [0, rows_pad): w rounded to tf32
    out[total + i] = v - hi;     // rows [rows_pad, 2 rows_pad): the remainder (exact), second operand of the three-pass mode
  }
}

// mode 2: the four parity classes (a, b) of the data gradient of a 3x3 STRIDE-2 convolution (conv_tc.cu MODE 3), one pair of
// hi / lo matrices [ci_pad][nt * Cout] per class, classes in the order (0,0), (0,1), (1,0), (1,1) with nt = 1, 2, 2, 4 taps:
// tap (ir, ic) of class (a, b) is filter element (kh, kw) = (a ? 2 ir : 1, b ? 2 ic : 1).
__global__ void pack_weight_tc_sub_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int rows_pad) {
  const int64_t total = (int64_t)rows_pad * 9 * Cout;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int co = (int)(i % Cout);
    const int q = (int)((i / Cout) % 9);           // tap slot over all classes
    const int r = (int)(i / ((int64_t)Cout * 9));   // input channel (row of the matrix)
    const int cls = q < 1 ? 0 : (q < 3 ? 1 : (q < 5 ? 2 : 3));
    const int before = cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? 3 : 5));
    const int a = cls >> 1, b = cls & 1, t = q - before, nt = (a ? 2 : 1) * (b ? 2 : 1);
    const int ir = b ? t >> 1 : t, ic = b ? t & 1 : 0;
    const int kh = a ? 2 * ir : 1, kw = b ? 2 * ic : 1;
    float v = 0.f;
    if (r < Cin) v = w[(((int64_t)co * Cin + r) * 3 + kh) * 3 + kw];
    uint32_t tt;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tt) : "f"(v));
    const float hi = __uint_as_float(tt);
    float* base = out + 2 * (int64_t)rows_pad * Cout * before;       // this class's hi matrix; lo follows it
    const int64_t e = ((int64_t)r * nt + t) * Cout + co;
    base[e] = hi;
    base[(int64_t)rows_pad * nt * Cout + e] = v - hi;
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient.  CTA = 128 threads = 2 pixel lanes x (8 ci-groups x 8 co-groups); each thread owns a
// 4(ci) x 4(co) x taps register tile and walks its lane's pixels of the staged tile.
constexpr int WG_CI = 32, WG_CO = 32;     // channels per CTA
constexpr int WG_TH = 4, WG_TW = 16;      // output pixels per staged tile

template <int KS>
__global__ void __launch_bounds__(128, 2)
wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int N, int H, int W,
             int Cin, int Ho, int Wo, int Cout, int S, int pad, int tiles_x, int tiles_per_img, int64_t total_tiles) {
  extern __shared__ __align__(16) float sm[];
  const int IH = (WG_TH - 1) * S + KS, IW = (WG_TW - 1) * S + KS;
  float* sx = sm;                              // [IH][IW][WG_CI]
  float* sd = sm + IH * IW * WG_CI;            // [WG_TH*WG_TW][WG_CO]
  const int t = threadIdx.x;
  const int lane = t >> 6;                     // pixel lane 0/1
  const int cig = (t & 7), cog = (t >> 3) & 7;
  const int ci0 = blockIdx.y * WG_CI, co0 = blockIdx.z * WG_CO;

  float acc[KS * KS][4][4];
#pragma unroll
  for (int a = 0; a < KS * KS; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][i][j] = 0.f;

  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = (int)(tile / tiles_per_img);
    const int tr = (int)(tile % tiles_per_img);
    const int oy0 = (tr / tiles_x) * WG_TH, ox0 = (tr % tiles_x) * WG_TW;
    const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
    __syncthreads();
    // stage X halo tile, 4 channels per thread-iteration
    for (int i = t; i < IH * IW * (WG_CI / 4); i += 128) {
      const int c4 = i % (WG_CI / 4), p = i / (WG_CI / 4);
      const int yy = p / IW, xx = p % IW;
      const int gy = iy0 + yy, gx = ix0 + xx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const float* src = x + (((int64_t)n * H + gy) * W + gx) * Cin;
        const int c = ci0 + 4 * c4;
        if (c + 3 < Cin && (Cin & 3) == 0) {
          v = __ldg(reinterpret_cast<const float4*>(src + c));
        } else {
          if (c < Cin) v.x = __ldg(src + c);
          if (c + 1 < Cin) v.y = __ldg(src + c + 1);
          if (c + 2 < Cin) v.z = __ldg(src + c + 2);
          if (c + 3 < Cin) v.w = __ldg(src + c + 3);
        }
      }
      *reinterpret_cast<float4*>(sx + p * WG_CI + 4 * c4) = v;
    }
    for (int i = t; i < WG_TH * WG_TW * (WG_CO / 4); i += 128) {
      const int c4 = i % (WG_CO / 4), p = i / (WG_CO / 4);
      const int oy = oy0 + p / WG_TW, ox = ox0 + p % WG_TW;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oy < Ho && ox < Wo) {
        const float* src = dy + (((int64_t)n * Ho + oy) * Wo + ox) * Cout;
        const int c = co0 + 4 * c4;
        if (c + 3 < Cout && (Cout & 3) == 0) {
          v = __ldg(reinterpret_cast<const float4*>(src + c));
        } else {
          if (c < Cout) v.x = __ldg(src + c);
          if (c + 1 < Cout) v.y = __ldg(src + c + 1);
          if (c + 2 < Cout) v.z = __ldg(src + c + 2);
          if (c + 3 < Cout) v.w = __ldg(src + c + 3);
        }
      }
      *reinterpret_cast<float4*>(sd + p * WG_CO + 4 * c4) = v;
    }
    __syncthreads();
    for (int p = lane; p < WG_TH * WG_TW; p += 2) {
      const int py = p / WG_TW, px = p % WG_TW;
      const float4 d = *reinterpret_cast<const float4*>(sd + p * WG_CO + 4 * cog);
      const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int kh = 0; kh < KS; ++kh)
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          const float4 xv =
              *reinterpret_cast<const float4*>(sx + ((py * S + kh) * IW + (px * S + kw)) * WG_CI + 4 * cig);
          const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[kh * KS + kw][i][j] = fmaf(xx[i], dd[j], acc[kh * KS + kw][i][j]);
        }
    }
  }
  // partial sums (OIHW) of this CTA's pixel share and lane -> slot (2 * blockIdx.x + lane); wgrad_sum_kernel adds the
  // slots in index order, so dW does not depend on which CTA finishes first (no floating-point atomics)
  float* mine = part + (int64_t)(2 * blockIdx.x + lane) * ((int64_t)Cout * Cin * (KS * KS));
#pragma unroll
  for (int a = 0; a < KS * KS; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = ci0 + 4 * cig + i, co = co0 + 4 * cog + j;
        if (ci < Cin && co < Cout) mine[((int64_t)co * Cin + ci) * (KS * KS) + a] = acc[a][i][j];
      }
}

// Weight gradient of a 3x3 convolution with a handful of input channels (the stem's first layer, Cin = 3: too narrow for
// TMA / the tensor-core tiling, and the register-tile kernel above spends its time staging 32-channel tiles that are 90 %
// padding - 2.8 ms per step).  Lane = output channel, warp = a slab of output rows: a thread keeps all 9 x CI products of its
// channel in registers, reads dY once (coalesced across the warp) and the 3 x 3 x CI input patch as warp-wide broadcasts.
// Per-row fp32 sums are folded into double; every warp writes its partial to its own slot, slots are added in order afterwards.
// XNCHW: x is the NCHW image as the caller holds it (plane stride H*W) instead of an NHWC copy
template <int CI, bool XNCHW = false>
__global__ void __launch_bounds__(256)
wgrad_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int N, int H, int W,
                       int Ho, int Wo, int Cout, int S, int nseg) {
  __shared__ float red[8][9 * CI][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.y * 32 + lane;
  const bool co_ok = co < Cout;
  constexpr int SEG = 64;                                 // output columns per work item
  const int64_t items = (int64_t)N * Ho * nseg;
  double dacc[9 * CI];
#pragma unroll
  for (int i = 0; i < 9 * CI; ++i) dacc[i] = 0.0;
  // work item = (output row, 64-column segment); items are dealt round-robin to the grid's warps - the order in which a
  // warp meets its items and the slot order of the final sum are fixed, so the result is reproducible
  for (int64_t it = (int64_t)blockIdx.x * 8 + warp; it < items; it += (int64_t)gridDim.x * 8) {
    const int64_t r = it / nseg;
    const int seg = (int)(it % nseg);
    const int n = (int)(r / Ho), oy = (int)(r % Ho);
    const float* dr = dy + r * Wo * Cout + co;
    const float* xrow[3];
    float rmask[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * S + kh - 1;
      const bool ok = iy >= 0 && iy < H;
      rmask[kh] = ok ? 1.f : 0.f;
      xrow[kh] = XNCHW ? x + ((int64_t)n * CI * H + (ok ? iy : 0)) * W : x + ((int64_t)n * H + (ok ? iy : 0)) * W * CI;
    }
    const int64_t plane = (int64_t)H * W;
    float acc[9 * CI];
#pragma unroll
    for (int i = 0; i < 9 * CI; ++i) acc[i] = 0.f;
    const int ox1 = min(Wo, (seg + 1) * SEG);
#pragma unroll 2
    for (int ox = seg * SEG; ox < ox1; ++ox) {
      const float d = co_ok ? __ldg(dr + (int64_t)ox * Cout) : 0.f;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * S + kw - 1;
        const bool ok = (unsigned)ix < (unsigned)W;
        const int ixc = ok ? ix : 0;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float dm = ok ? d * rmask[kh] : 0.f;
#pragma unroll
          for (int c = 0; c < CI; ++c)
            acc[(kh * 3 + kw) * CI + c] = fmaf(__ldg(XNCHW ? xrow[kh] + c * plane + ixc : xrow[kh] + ixc * CI + c), dm,
                                               acc[(kh * 3 + kw) * CI + c]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 9 * CI; ++i) dacc[i] += acc[i];
  }
#pragma unroll
  for (int i = 0; i < 9 * CI; ++i) red[warp][i][lane] = (float)dacc[i];
  __syncthreads();
  // block partial = the eight warps in warp order -> slot blockIdx.x
  float* mine = part + (int64_t)blockIdx.x * ((int64_t)Cout * CI * 9);
  for (int e = threadIdx.x; e < 9 * CI * 32; e += 256) {
    const int i = e >> 5, l = e & 31;
    const int c_o = blockIdx.y * 32 + l;
    if (c_o >= Cout) continue;
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sum += red[wv][i][l];
    const int t = i / CI, c = i % CI;
    mine[((int64_t)c_o * CI + c) * 9 + t] = sum;
  }
}

// Forward of a 3x3 convolution with a handful of input channels read straight from the NCHW image (the stem's first layer):
// thread = output pixel, all Cout channels in registers (two halves of 16), the 3 x 3 x CI patch loaded once per thread
// (coalesced along x within each plane), weights [tap][ci][co] in shared memory.  Replaces an NCHW -> NHWC copy of the image
// (0.59 ms per step) plus the generic register-tile kernel (0.81 ms); fp32 FMA chains, tap-major like conv_direct.
template <int CI>
__global__ void __launch_bounds__(128)
stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_oihw, const float* __restrict__ bias,
                float* __restrict__ y, int N, int H, int W, int Ho, int Wo, int Cout, int S) {
  extern __shared__ float sw[];                 // [9][CI][Cout]
  for (int i = threadIdx.x; i < 9 * CI * Cout; i += 128) {
    const int co = i % Cout, ci = (i / Cout) % CI, t = i / (Cout * CI);
    sw[i] = w_oihw[((int64_t)co * CI + ci) * 9 + t];
  }
  __syncthreads();
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t plane = (int64_t)H * W;
  for (int64_t p = blockIdx.x * 128ll + threadIdx.x; p < total; p += (int64_t)gridDim.x * 128) {
    const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((int64_t)Wo * Ho));
    float xv[9 * CI];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iy = oy * S + kh - 1, ix = ox * S + kw - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int c = 0; c < CI; ++c)
          xv[(kh * 3 + kw) * CI + c] = ok ? __ldg(x + ((int64_t)n * CI + c) * plane + (int64_t)iy * W + ix) : 0.f;
      }
    float* yo = y + p * Cout;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = (bias && c0 + j < Cout) ? bias[c0 + j] : 0.f;
#pragma unroll
      for (int t = 0; t < 9 * CI; ++t) {
        const float4* wr = reinterpret_cast<const float4*>(sw + t * Cout + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q + 0] = fmaf(xv[t], wv.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(xv[t], wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv[t], wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv[t], wv.w, acc[4 * q + 3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(yo + c0 + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
  }
}

__global__ void __launch_bounds__(kT) wgrad_sum_kernel(const float* __restrict__ part, float* __restrict__ dw, int64_t n,
                                                       int slots) {
  const int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < slots; ++k) s += part[(int64_t)k * n + i];
  dw[i] = s;
}

// ---------------------------------------------------------------------------------------------
// per-channel reductions over [npix][C] (C innermost).  Threads g < R*C own channel g % C and rows
// g / C, g / C + R, ...; the threads of a block that share a channel are summed by ONE of them in thread order, every
// block writes its [2][C] partial to `part[blockIdx.x]`, and chan_sum_kernel adds the blocks in index order: the
// result is bit-reproducible (no atomics).
// kind 0: (sum x, sum x^2)            -> BN batch statistics
// kind 1: (sum dz, sum dz * xhat)     -> BN backward, dz = dout * (out > 0 if relu)
// kind 2: (sum x, -)                  -> bias gradient
constexpr int kMaxRedBlocks = 148 * 4;
template <int KIND>
__global__ void __launch_bounds__(kT)
chan_reduce_kernel(const float* __restrict__ a, const float* __restrict__ y, const float* __restrict__ out,
                   const float* __restrict__ mean, const float* __restrict__ invstd, int relu, int64_t npix, int C,
                   double* __restrict__ part /* [gridDim.x][2][C] */) {
  extern __shared__ double sacc[];   // [2][C]
  __shared__ double sd0[kT], sd1[kT];
  for (int i = threadIdx.x; i < 2 * C; i += kT) sacc[i] = 0.0;
  const int64_t T = (int64_t)gridDim.x * kT;
  const int64_t R = T / C;
  const int64_t g = blockIdx.x * (int64_t)kT + threadIdx.x;
  double d0 = 0.0, d1 = 0.0;
  if (R > 0 && g < R * C) {
    const int c = (int)(g % C);
    float s0 = 0.f, s1 = 0.f;
    float m = 0.f, is = 0.f;
    if (KIND == 1) { m = mean[c]; is = invstd[c]; }
    int cnt = 0;
    for (int64_t r = g / C; r < npix; r += R) {
      const int64_t i = r * C + c;
      if (KIND == 0) {
        const float v = a[i];
        s0 += v; s1 = fmaf(v, v, s1);
      } else if (KIND == 1) {
        float dz = a[i];
        if (relu && !(out[i] > 0.f)) dz = 0.f;
        s0 += dz; d1 += (double)dz * ((double)(y[i] - m) * (double)is);
      } else {
        s0 += a[i];
      }
      if (++cnt == 256) { d0 += s0; d1 += s1; s0 = s1 = 0.f; cnt = 0; }   // bound the fp32 run length
    }
    d0 += s0; d1 += s1;
  }
  sd0[threadIdx.x] = d0; sd1[threadIdx.x] = d1;
  __syncthreads();
  // thread t < min(C, kT) is the first thread of its channel in this block: it adds the block's threads t, t + C, ...
  if (threadIdx.x < C) {
    double t0 = 0.0, t1 = 0.0;
    for (int j = threadIdx.x; j < kT; j += C) { t0 += sd0[j]; t1 += sd1[j]; }
    const int c = (int)(g % C);
    sacc[c] = t0; sacc[C + c] = t1;
  }
  __syncthreads();
  double* mine = part + (int64_t)blockIdx.x * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += kT) mine[i] = sacc[i];
}

// The same reductions with thread = FOUR consecutive channels (C % 4 == 0, 16-byte aligned rows, < 2^31 elements): 128-bit
// loads and 32-bit index arithmetic (the scalar kernels above spend most of their issue slots on 64-bit div / mod).
template <int KIND>
__global__ void __launch_bounds__(kT)
chan_reduce4_kernel(const float4* __restrict__ a, const float4* __restrict__ y, const float4* __restrict__ out,
                    const float* __restrict__ mean, const float* __restrict__ invstd, int relu, uint32_t npix, uint32_t C4,
                    double* __restrict__ part /* [gridDim.x][2][C] */) {
  extern __shared__ double sacc[];   // [2][C]
  __shared__ double sd[2][4][kT];
  const uint32_t C = C4 * 4;
  const uint32_t T = gridDim.x * kT;
  const uint32_t R = T / C4;
  const uint32_t g = blockIdx.x * kT + threadIdx.x;
  double d0[4] = {0.0, 0.0, 0.0, 0.0}, d1[4] = {0.0, 0.0, 0.0, 0.0};
  if (R > 0 && g < R * C4) {
    const uint32_t c4 = g % C4;
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), is = m;
    if (KIND == 1) { m = *reinterpret_cast<const float4*>(mean + 4 * c4); is = *reinterpret_cast<const float4*>(invstd + 4 * c4); }
    int cnt = 0;
#pragma unroll 4                                 // four rows' loads in flight; the accumulation order is unchanged
    for (uint32_t r = g / C4; r < npix; r += R) {
      const uint32_t i = r * C4 + c4;
      const float4 v = __ldg(a + i);
      if (KIND == 0) {
        s0[0] += v.x; s0[1] += v.y; s0[2] += v.z; s0[3] += v.w;
        s1[0] = fmaf(v.x, v.x, s1[0]); s1[1] = fmaf(v.y, v.y, s1[1]); s1[2] = fmaf(v.z, v.z, s1[2]); s1[3] = fmaf(v.w, v.w, s1[3]);
      } else if (KIND == 1) {
        float4 dz = v;
        if (relu) {
          const float4 o = __ldg(out + i);
          if (!(o.x > 0.f)) dz.x = 0.f;
          if (!(o.y > 0.f)) dz.y = 0.f;
          if (!(o.z > 0.f)) dz.z = 0.f;
          if (!(o.w > 0.f)) dz.w = 0.f;
        }
        const float4 yy = __ldg(y + i);
        s0[0] += dz.x; s0[1] += dz.y; s0[2] += dz.z; s0[3] += dz.w;
        // sum dz * xhat straight into double: this sum and sum dz are what BatchNorm's backward subtracts from dz (a
        // cancellation that amplifies their rounding by the network's conditioning); the kernel is memory-bound either way
        d1[0] += (double)dz.x * ((double)(yy.x - m.x) * (double)is.x); d1[1] += (double)dz.y * ((double)(yy.y - m.y) * (double)is.y);
        d1[2] += (double)dz.z * ((double)(yy.z - m.z) * (double)is.z); d1[3] += (double)dz.w * ((double)(yy.w - m.w) * (double)is.w);
      } else {
        s0[0] += v.x; s0[1] += v.y; s0[2] += v.z; s0[3] += v.w;
      }
      if (++cnt == 64) {                        // bound the fp32 run length
#pragma unroll
        for (int j = 0; j < 4; ++j) { d0[j] += s0[j]; d1[j] += s1[j]; s0[j] = s1[j] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { d0[j] += s0[j]; d1[j] += s1[j]; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { sd[0][j][threadIdx.x] = d0[j]; sd[1][j][threadIdx.x] = d1[j]; }
  for (uint32_t i = threadIdx.x; i < 2 * C; i += kT) sacc[i] = 0.0;
  __syncthreads();
  // thread t < min(C4, kT) is the first thread of its channel group in this block: it adds threads t, t + C4, ... in order
  if (threadIdx.x < C4) {
    const uint32_t c4 = g % C4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double t0 = 0.0, t1 = 0.0;
      for (uint32_t k = threadIdx.x; k < kT; k += C4) { t0 += sd[0][j][k]; t1 += sd[1][j][k]; }
      sacc[4 * c4 + j] = t0; sacc[C + 4 * c4 + j] = t1;
    }
  }
  __syncthreads();
  double* mine = part + (int64_t)blockIdx.x * 2 * C;
  for (uint32_t i = threadIdx.x; i < 2 * C; i += kT) mine[i] = sacc[i];
}

// acc[i] = sum over the blocks' partials, one WARP per output: lane l adds blocks l, l + 32, ... in index order and the 32
// lane sums are combined by a fixed shuffle tree - the order depends only on (nblocks), never on scheduling.
// (round 2 measured the one-thread-per-output version at 57 us per launch, 7 ms per training step: 592 dependent loads)
// dup_off > 0: the sums are also written at acc[dup_off + i] (SyncBN backward: the copy that gets all-reduced);
// tail_idx >= 0: acc[tail_idx] = tail_val (SyncBN forward: this rank's sample count, all-reduced with the sums)
__global__ void __launch_bounds__(kT) chan_sum_kernel(const double* __restrict__ part, int nblocks, int n2c,
                                                      double* __restrict__ acc, int dup_off, int tail_idx, double tail_val) {
  const int i = (blockIdx.x * kT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n2c) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += part[(int64_t)b * n2c + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) {
    acc[i] = s;
    if (dup_off > 0) acc[dup_off + i] = s;
    if (i == 0 && tail_idx >= 0) acc[tail_idx] = tail_val;
  }
}

__global__ void move_double_kernel(double* dst, const double* src) { *dst = *src; }

// npix < 0: the sample count is read from acc[2C] on the device (SyncBN: the all-reduced count, no host round trip)
__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, double npix, float eps, float momentum,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (npix < 0.0) npix = acc[2 * C];
  const double m = acc[c] / npix;
  double var = acc[C + c] / npix - m * m;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)m;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = npix > 1.0 ? var * npix / (npix - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// out = [relu]((y - mean) * invstd * gamma + beta [+ res])
// IT: index type - uint32_t when the tensor has < 2^32 elements (a 64-bit modulo per element is most of this kernel's time)
template <typename IT>
__global__ void __launch_bounds__(kT)
bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ res, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                float* __restrict__ out, int relu, int64_t total_, int C_, int rnd) {
  const IT total = (IT)total_, C = (IT)C_;
  for (IT i = blockIdx.x * (IT)kT + threadIdx.x; i < total; i += (IT)gridDim.x * kT) {
    const int c = (int)(i % C);
    float v = fmaf((y[i] - mean[c]) * invstd[c], gamma[c], beta[c]);
    if (res) v += res[i];
    out[i] = rt(relu ? fmaxf(v, 0.f) : v, rnd);
  }
}

// dy = gamma * invstd * (dz - dbeta/M - xhat * dgamma/M); dres = dz
template <typename IT>
__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ out,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const double* __restrict__ acc, const double* __restrict__ acc_local, float* __restrict__ dy,
                    float* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int relu, int64_t total_,
                    int C_, float invM, int rnd) {
  const IT total = (IT)total_, C = (IT)C_;
  if (invM < 0.f) invM = (float)(1.0 / acc_local[4 * C_]);      // SyncBN: global sample count kept on the device at workspace[4C]
  for (IT i = blockIdx.x * (IT)kT + threadIdx.x; i < total; i += (IT)gridDim.x * kT) {
    const int c = (int)(i % C);
    float dz = dout[i];
    if (relu && !(out[i] > 0.f)) dz = 0.f;
    const float is = invstd[c];
    const double xhat = (double)(y[i] - mean[c]) * (double)is;
    const double br = (double)dz - acc[c] * (double)invM - xhat * (acc[C + c] * (double)invM);
    dy[i] = rt((float)((double)gamma[c] * (double)is * br), rnd);
    if (dres) dres[i] = dz;
    if (i < C) { dbeta[c] = (float)acc_local[c]; dgamma[c] = (float)acc_local[C + c]; }   // parameter gradients stay rank-local (DDP averages them)
  }
}

// vectorised forms of the two kernels above (C % 4 == 0, aligned, < 2^31 elements)
__global__ void __launch_bounds__(kT)
bn_apply4_kernel(const float4* __restrict__ y, const float4* __restrict__ res, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                 float4* __restrict__ out, int relu, uint32_t total4, uint32_t C4, int rnd) {
  const float lo = relu ? 0.f : -INFINITY;
  for (uint32_t i = blockIdx.x * kT + threadIdx.x; i < total4; i += gridDim.x * kT) {
    const uint32_t c = (i % C4) * 4;
    const float4 v = __ldg(y + i);
    const float4 m = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = fmaf((v.x - m.x) * is.x, g.x, b.x); o.y = fmaf((v.y - m.y) * is.y, g.y, b.y);
    o.z = fmaf((v.z - m.z) * is.z, g.z, b.z); o.w = fmaf((v.w - m.w) * is.w, g.w, b.w);
    if (res) { const float4 r = __ldg(res + i); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    o.x = rt(fmaxf(o.x, lo), rnd); o.y = rt(fmaxf(o.y, lo), rnd); o.z = rt(fmaxf(o.z, lo), rnd); o.w = rt(fmaxf(o.w, lo), rnd);
    out[i] = o;
  }
}

__global__ void __launch_bounds__(kT)
bn_bwd_apply4_kernel(const float4* __restrict__ dout, const float4* __restrict__ y, const float4* __restrict__ out,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                     const double* __restrict__ acc, const double* __restrict__ acc_local, float4* __restrict__ dy,
                     float4* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int relu,
                     uint32_t total4, uint32_t C4, float invM, int rnd) {
  const uint32_t C = C4 * 4;
  if (invM < 0.f) invM = (float)(1.0 / acc_local[4 * C]);       // SyncBN: global sample count kept on the device at workspace[4C]
  for (uint32_t i = blockIdx.x * kT + threadIdx.x; i < total4; i += gridDim.x * kT) {
    const uint32_t c = (i % C4) * 4;
    float4 dz = __ldg(dout + i);
    if (relu) {
      const float4 o = __ldg(out + i);
      if (!(o.x > 0.f)) dz.x = 0.f;
      if (!(o.y > 0.f)) dz.y = 0.f;
      if (!(o.z > 0.f)) dz.z = 0.f;
      if (!(o.w > 0.f)) dz.w = 0.f;
    }
    const float4 yy = __ldg(y + i);
    const float4 m = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float dzv[4] = {dz.x, dz.y, dz.z, dz.w}, yv[4] = {yy.x, yy.y, yy.z, yy.w};
    const float mv[4] = {m.x, m.y, m.z, m.w}, isv[4] = {is.x, is.y, is.z, is.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // dz - mean(dz) - xhat * mean(dz * xhat) in double: the subtraction cancels most of dz in the deep layers
      const double xhat = (double)(yv[j] - mv[j]) * (double)isv[j];
      const double br = (double)dzv[j] - acc[c + j] * (double)invM - xhat * (acc[C + c + j] * (double)invM);
      o[j] = rt((float)((double)gv[j] * (double)isv[j] * br), rnd);
    }
    dy[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (dres) dres[i] = dz;
    if (i < C4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { dbeta[c + j] = (float)acc_local[c + j]; dgamma[c + j] = (float)acc_local[C + c + j]; }
    }
  }
}

// C % 4 != 0 (the K-class head BatchNorms, C = 19): the tensor is still a dense array of floats, so it is walked as float4s
// of the FLAT index; every component looks its channel up on its own ((4 i + j) % C).  Needs total % 4 == 0.
__global__ void __launch_bounds__(kT)
bn_apply_flat4_kernel(const float4* __restrict__ y, const float4* __restrict__ res, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                      float4* __restrict__ out, int relu, uint32_t total4, uint32_t C, int rnd) {
  const float lo = relu ? 0.f : -INFINITY;
  for (uint32_t i = blockIdx.x * kT + threadIdx.x; i < total4; i += gridDim.x * kT) {
    const float4 v = __ldg(y + i);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) r = __ldg(res + i);
    const float vv[4] = {v.x, v.y, v.z, v.w}, rr[4] = {r.x, r.y, r.z, r.w};
    float o[4];
    uint32_t c = (uint32_t)(((uint64_t)i * 4) % C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = fmaf((vv[j] - __ldg(mean + c)) * __ldg(invstd + c), __ldg(gamma + c), __ldg(beta + c));
      if (res) t += rr[j];
      o[j] = rt(fmaxf(t, lo), rnd);
      if (++c == C) c = 0;
    }
    out[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(kT)
bn_bwd_apply_flat4_kernel(const float4* __restrict__ dout, const float4* __restrict__ y, const float4* __restrict__ out,
                          const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                          const double* __restrict__ acc, const double* __restrict__ acc_local, float4* __restrict__ dy,
                          float4* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int relu,
                          uint32_t total4, uint32_t C, float invM, int rnd) {
  if (invM < 0.f) invM = (float)(1.0 / acc_local[4 * C]);
  for (uint32_t i = blockIdx.x * kT + threadIdx.x; i < total4; i += gridDim.x * kT) {
    float4 dz = __ldg(dout + i);
    if (relu) {
      const float4 o = __ldg(out + i);
      if (!(o.x > 0.f)) dz.x = 0.f;
      if (!(o.y > 0.f)) dz.y = 0.f;
      if (!(o.z > 0.f)) dz.z = 0.f;
      if (!(o.w > 0.f)) dz.w = 0.f;
    }
    const float4 yy = __ldg(y + i);
    const float dzv[4] = {dz.x, dz.y, dz.z, dz.w}, yv[4] = {yy.x, yy.y, yy.z, yy.w};
    float o[4];
    uint32_t c = (uint32_t)(((uint64_t)i * 4) % C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float is = __ldg(invstd + c);
      const double xhat = (double)(yv[j] - __ldg(mean + c)) * (double)is;
      const double br = (double)dzv[j] - acc[c] * (double)invM - xhat * (acc[C + c] * (double)invM);
      o[j] = rt((float)((double)__ldg(gamma + c) * (double)is * br), rnd);
      if (++c == C) c = 0;
    }
    dy[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (dres) dres[i] = dz;
    if (i * 4 < C) {
      for (uint32_t j = i * 4; j < i * 4 + 4 && j < C; ++j) { dbeta[j] = (float)acc_local[j]; dgamma[j] = (float)acc_local[C + j]; }
    }
  }
}

__global__ void acc_to_float_kernel(const double* __restrict__ acc, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[c];
}

// ---------------------------------------------------------------------------------------------
// All-reduce (sum) of a short double vector over the GPUs of one NVSwitch box through PEER MEMORY, for the SyncBN statistics
// (configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:20): 2C + 1 doubles per layer and direction, 122 times per training
// step.  NCCL spends ~50 us per such message (launch + protocol latency, 6.7 ms per step at N = 8); here ONE launch per
// rank stores the rank's vector into its slot of EVERY peer's buffer (NVLink P2P stores), releases a per-source flag on each
// peer, waits for the flags of all sources in its OWN buffer and adds the slots in rank order - so every rank gets the same
// bits, whatever the arrival order.  Slots and flags are double-buffered by the parity of a sequence number that all ranks
// advance in lock step: a peer signals collective s + 1 only after it has finished reading collective s.
// Buffer layout on every rank (symmetric allocation): data[2][kPeerMax][nmax] doubles, then flags[2][kPeerMax] uint32.
constexpr int kPeerMax = 16;
struct PeerPtrs { unsigned long long p[kPeerMax]; };

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const double* __restrict__ local, int n, PeerPtrs peers, int rank, int world, unsigned seq, int nmax,
                      double* __restrict__ out) {
  const int par = (int)(seq & 1u);
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(peers.p[p]) + ((size_t)par * kPeerMax + rank) * nmax;
    for (int i = threadIdx.x; i < n; i += 256) dst[i] = local[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned* theirs = reinterpret_cast<unsigned*>(reinterpret_cast<double*>(peers.p[threadIdx.x]) + (size_t)2 * kPeerMax * nmax) +
                       par * kPeerMax + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(seq) : "memory");
    const unsigned* mine = reinterpret_cast<const unsigned*>(reinterpret_cast<const double*>(peers.p[rank]) + (size_t)2 * kPeerMax * nmax) +
                           par * kPeerMax + threadIdx.x;
    unsigned v = 0, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (++spins > (1u << 30)) __trap();          // a peer that never arrives must not hang the GPU
    } while (v != seq);
  }
  __syncthreads();
  const double* src = reinterpret_cast<const double*>(peers.p[rank]) + (size_t)par * kPeerMax * nmax;
  for (int i = threadIdx.x; i < n; i += 256) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += src[(size_t)p * nmax + i];
    out[i] = s;
  }
}

// workspace layout of the per-channel reductions (doubles): [0, 4C + 8) accumulators (2C of a reduction; SyncBN keeps a second
// copy and the sample count there), then kMaxRedBlocks block partials of 2C each
inline bool vec4_ok(int C, int64_t total, const void* p0, const void* p1 = nullptr, const void* p2 = nullptr,
                    const void* p3 = nullptr, const void* p4 = nullptr) {
  return C % 4 == 0 && total < (1ll << 31) &&
         ((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)p3 | (uintptr_t)p4) % 16 == 0;
}

inline int64_t bn_ws_doubles(int C) { return 4 * (int64_t)C + 8 + (int64_t)kMaxRedBlocks * 2 * C; }

template <int KIND>
int chan_reduce(const float* a, const float* y, const float* out, const float* mean, const float* invstd, int relu,
                int64_t npix, int C, double* ws, cudaStream_t st, int dup_off = 0, int tail_idx = -1, double tail_val = 0.0) {
  double* part = ws + 4 * (int64_t)C + 8;
  const int grid = grid1d(npix * C, 4);                     // <= kMaxRedBlocks
  if (vec4_ok(C, npix * C, a, y, out) && ((uintptr_t)mean | (uintptr_t)invstd) % 16 == 0)
    chan_reduce4_kernel<KIND><<<grid, kT, sizeof(double) * 2 * C, st>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(out), mean, invstd,
        relu, (uint32_t)npix, (uint32_t)(C / 4), part);
  else
  chan_reduce_kernel<KIND><<<grid, kT, sizeof(double) * 2 * C, st>>>(a, y, out, mean, invstd, relu, npix, C, part);
  chan_sum_kernel<<<ceil_div(2 * C * 32, kT), kT, 0, st>>>(part, grid, 2 * C, ws, dup_off, tail_idx, tail_val);
  LEDB_LAUNCH_OK("chan_reduce_kernel");
  return LEDB200_OK;
}

// ---------------------------------------------------------------------------------------------
// bilinear resize (align_corners=False), generic channel count
__global__ void __launch_bounds__(kT)
resize_fwd_kernel(const float* __restrict__ src, float* __restrict__ out, int N, int h, int w, int H, int W, int C,
                  float sh, float sw, int rnd) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
    bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
    const float* s = src + (int64_t)n * h * w * C + c;
    const float r0 = fmaf(s[((int64_t)y0 * w + x1) * C], lx1, s[((int64_t)y0 * w + x0) * C] * lx0);
    const float r1 = fmaf(s[((int64_t)y1 * w + x1) * C], lx1, s[((int64_t)y1 * w + x0) * C] * lx0);
    out[i] = rt(fmaf(r1, ly1, r0 * ly0), rnd);
  }
}

// candidate outputs that may read source index `s`: the source coordinate of output d is scale * (d + 0.5) - 0.5 and
// it reads floor(.) and floor(.) + 1, so d lies in ((s - 0.5) / scale - 0.5, (s + 1.5) / scale - 0.5); one pixel of slack
// either side, the exact membership is re-tested per candidate
__device__ __forceinline__ void gather_range(int s, float scale, int out_size, int& lo, int& hi) {
  lo = (int)floorf(((float)s - 0.5f) / scale - 0.5f) - 1;
  hi = (int)ceilf(((float)s + 1.5f) / scale - 0.5f) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

// dsrc[n, ys, xs, c] = sum over the outputs (y, x) whose bilinear footprint holds (ys, xs) of dout * weight, rows then
// columns in ascending order: the transpose of resize_fwd_kernel as a GATHER (no atomics, fixed summation order)
__global__ void __launch_bounds__(kT)
resize_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dsrc, int N, int h, int w, int H, int W, int C,
                  float sh, float sw, int rnd) {
  const int64_t total = (int64_t)N * h * w * C;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int xs = (int)(p % w), ys = (int)((p / w) % h), n = (int)(p / ((int64_t)w * h));
    int ylo, yhi, xlo, xhi;
    gather_range(ys, sh, H, ylo, yhi);
    gather_range(xs, sw, W, xlo, xhi);
    const float* g = dout + (int64_t)n * H * W * C + c;
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      int y0, y1;
      float ly0, ly1;
      bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
      float wy = 0.f;
      if (y0 == ys) wy += ly0;
      if (y1 == ys) wy += ly1;                     // y0 == y1 at the clamped border: both weights land on the same pixel
      if (y0 != ys && y1 != ys) continue;
      float row = 0.f;
      for (int x = xlo; x <= xhi; ++x) {
        int x0, x1;
        float lx0, lx1;
        bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
        if (x0 != xs && x1 != xs) continue;
        float wx = 0.f;
        if (x0 == xs) wx += lx0;
        if (x1 == xs) wx += lx1;
        row = fmaf(g[((int64_t)y * W + x) * C], wx, row);
      }
      acc = fmaf(row, wy, acc);
    }
    dsrc[i] = rt(acc, rnd);
  }
}

// Row-decomposed forms of the two resize kernels (same arithmetic, same summation order, bit-identical results): one block
// row per (image, output row), so the vertical coordinates are block-uniform and the per-element index arithmetic is one
// 32-bit div / mod by C; V = 4 channels per thread through 128-bit accesses when C % 4 == 0.
// (round 2 measured the flat kernels at 10x the HBM time on the head's logit ladder: 64-bit div / mod per element and, in
// the backward gather, a bilinear_coord per (row, column) candidate pair instead of per row and per column)
template <int V> struct VecT { using type = float; };
template <> struct VecT<4> { using type = float4; };
__device__ __forceinline__ float vget(const float& v, int) { return v; }
__device__ __forceinline__ float vget(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
__device__ __forceinline__ void vset(float& v, int, float x) { v = x; }
__device__ __forceinline__ void vset(float4& v, int j, float x) { if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x; }

template <int V>
__global__ void __launch_bounds__(kT)
resize_fwd_row_kernel(const float* __restrict__ src, float* __restrict__ out, int h, int w, int H, int W, int C, float sh,
                      float sw, int rnd) {
  using T = typename VecT<V>::type;
  const uint32_t row = blockIdx.x, n = row / (uint32_t)H, y = row % (uint32_t)H;
  int y0, y1;
  float ly0, ly1;
  bilinear_coord((int)y, sh, h, y0, y1, ly0, ly1);
  const float* s0 = src + ((int64_t)n * h + y0) * w * C;
  const float* s1 = src + ((int64_t)n * h + y1) * w * C;
  float* o = out + (int64_t)row * W * C;
  const uint32_t len = (uint32_t)W * C;
  for (uint32_t j = (blockIdx.y * kT + threadIdx.x) * V; j < len; j += gridDim.y * kT * V) {
    const uint32_t x = j / (uint32_t)C, c = j % (uint32_t)C;
    int x0, x1;
    float lx0, lx1;
    bilinear_coord((int)x, sw, w, x0, x1, lx0, lx1);
    const T a = __ldg(reinterpret_cast<const T*>(s0 + x0 * C + c)), b = __ldg(reinterpret_cast<const T*>(s0 + x1 * C + c));
    const T cc = __ldg(reinterpret_cast<const T*>(s1 + x0 * C + c)), d = __ldg(reinterpret_cast<const T*>(s1 + x1 * C + c));
    T r;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float r0 = fmaf(vget(b, k), lx1, vget(a, k) * lx0);
      const float r1 = fmaf(vget(d, k), lx1, vget(cc, k) * lx0);
      vset(r, k, rt(fmaf(r1, ly1, r0 * ly0), rnd));
    }
    *reinterpret_cast<T*>(o + j) = r;
  }
}

// The horizontal candidates and weights of a source column depend on the column only: a block (one source row) computes
// them ONCE per column into shared memory - the flat form recomputed up to eight bilinear_coord per (column, channel)
// element, 3/4 of its instructions at C = 19.  Columns with more than XMAX candidates (down-scaling) are walked on the fly.
constexpr int kBwdXMax = 8;
struct BwdCol { int xlo, nx; float wx[kBwdXMax]; };

template <int V>
__global__ void __launch_bounds__(kT)
resize_bwd_row_kernel(const float* __restrict__ dout, float* __restrict__ dsrc, int h, int w, int H, int W, int C, float sh,
                      float sw, int rnd) {
  using T = typename VecT<V>::type;
  extern __shared__ unsigned char bwd_smem[];
  BwdCol* cols = reinterpret_cast<BwdCol*>(bwd_smem);
  __shared__ int s_ylo, s_ny;
  __shared__ float s_wy[16];
  const uint32_t row = blockIdx.x, n = row / (uint32_t)h, ys = row % (uint32_t)h;
  for (int xs = threadIdx.x; xs < w; xs += kT) {
    int xlo, xhi;
    gather_range(xs, sw, W, xlo, xhi);
    BwdCol c;
    c.xlo = xlo; c.nx = xhi - xlo + 1;
#pragma unroll
    for (int k = 0; k < kBwdXMax; ++k) {
      float wx = -1.f;                                       // -1: not a member (a member's weight may be exactly 0)
      if (k < c.nx) {
        int x0, x1;
        float lx0, lx1;
        bilinear_coord(xlo + k, sw, w, x0, x1, lx0, lx1);
        if (x0 == xs || x1 == xs) {
          wx = 0.f;
          if (x0 == xs) wx += lx0;
          if (x1 == xs) wx += lx1;
        }
      }
      c.wx[k] = wx;
    }
    cols[xs] = c;
  }
  if (threadIdx.x == 0) {
    int ylo, yhi;
    gather_range((int)ys, sh, H, ylo, yhi);
    int ny = yhi - ylo + 1;
    if (ny > 16) ny = -1;                                    // too many vertical candidates for the table: computed per thread
    s_ylo = ylo; s_ny = ny;
    for (int k = 0; k < 16 && k < ny; ++k) {
      int y0, y1;
      float ly0, ly1;
      bilinear_coord(ylo + k, sh, h, y0, y1, ly0, ly1);
      float wy = -1.f;
      if (y0 == (int)ys || y1 == (int)ys) {
        wy = 0.f;
        if (y0 == (int)ys) wy += ly0;
        if (y1 == (int)ys) wy += ly1;
      }
      s_wy[k] = wy;
    }
  }
  __syncthreads();
  int ylo = s_ylo, yhi;
  const int ny = s_ny;
  if (ny < 0) gather_range((int)ys, sh, H, ylo, yhi); else yhi = ylo + ny - 1;
  const float* g = dout + (int64_t)n * H * W * C;
  float* o = dsrc + (int64_t)row * w * C;
  const uint32_t len = (uint32_t)w * C;
  for (uint32_t j = (blockIdx.y * kT + threadIdx.x) * V; j < len; j += gridDim.y * kT * V) {
    const uint32_t xs = j / (uint32_t)C, c = j % (uint32_t)C;
    const BwdCol& col = cols[xs];
    const int xlo = col.xlo, nx = col.nx;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      float wy;
      if (ny >= 0) {
        wy = s_wy[y - ylo];
        if (wy < 0.f) continue;
      } else {
        int y0, y1;
        float ly0, ly1;
        bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
        if (y0 != (int)ys && y1 != (int)ys) continue;
        wy = 0.f;
        if (y0 == (int)ys) wy += ly0;
        if (y1 == (int)ys) wy += ly1;
      }
      float rowv[V];
#pragma unroll
      for (int k = 0; k < V; ++k) rowv[k] = 0.f;
      const float* gr = g + ((int64_t)y * W) * C + c;
      if (nx <= kBwdXMax) {
#pragma unroll
        for (int k = 0; k < kBwdXMax; ++k) {
          if (k < nx && col.wx[k] >= 0.f) {
            const T v = __ldg(reinterpret_cast<const T*>(gr + (xlo + k) * C));
#pragma unroll
            for (int q = 0; q < V; ++q) rowv[q] = fmaf(vget(v, q), col.wx[k], rowv[q]);
          }
        }
      } else {
        for (int x = xlo; x < xlo + nx; ++x) {
          int x0, x1;
          float lx0, lx1;
          bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
          if (x0 != (int)xs && x1 != (int)xs) continue;
          float wx = 0.f;
          if (x0 == (int)xs) wx += lx0;
          if (x1 == (int)xs) wx += lx1;
          const T v = __ldg(reinterpret_cast<const T*>(gr + x * C));
#pragma unroll
          for (int q = 0; q < V; ++q) rowv[q] = fmaf(vget(v, q), wx, rowv[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] = fmaf(rowv[q], wy, acc[q]);
    }
    T r;
#pragma unroll
    for (int q = 0; q < V; ++q) vset(r, q, rt(acc[q], rnd));
    *reinterpret_cast<T*>(o + j) = r;
  }
}

// out = [relu](a [+ b]);  backward: dx = dout * (out > 0)
__global__ void __launch_bounds__(kT)
add_relu_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int relu,
                int64_t n, int rnd) {
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT) {
    float v = a[i];
    if (b) v += b[i];
    out[i] = rt(relu ? fmaxf(v, 0.f) : v, rnd);
  }
}
__global__ void __launch_bounds__(kT)
relu_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dx, int64_t n,
                int rnd) {
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT)
    dx[i] = out[i] > 0.f ? rt(dout[i], rnd) : 0.f;
}

// AvgPool2d(k,s,p,count_include_pad=True) (k > 0) or global average (k == 0)
__device__ __forceinline__ float pool_inv(int o, int k, int s, int p, int extent) {
  const int a0 = o * s - p;
  const int a1 = min(a0 + k, extent + p);
  return (float)(a1 - a0);
}
__global__ void __launch_bounds__(kT)
avgpool_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C, int Ho, int Wo,
                   int k, int s, int p, int rnd) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    int64_t q = i / C;
    const int ox = (int)(q % Wo), oy = (int)((q / Wo) % Ho), n = (int)(q / ((int64_t)Wo * Ho));
    int y0 = 0, y1 = H, x0 = 0, x1 = W;
    float inv = 1.f / (float)(H * W);
    if (k > 0) {
      inv = 1.f / (pool_inv(oy, k, s, p, H) * pool_inv(ox, k, s, p, W));
      y0 = max(oy * s - p, 0); y1 = min(oy * s - p + k, H);
      x0 = max(ox * s - p, 0); x1 = min(ox * s - p + k, W);
    }
    float acc = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) acc += in[(((int64_t)n * H + y) * W + x) * C + c];
    out[i] = rt(acc * inv, rnd);
  }
}
__global__ void __launch_bounds__(kT)
avgpool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int N, int H, int W, int C, int Ho, int Wo,
                   int k, int s, int p, int rnd) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    int64_t q = i / C;
    const int x = (int)(q % W), y = (int)((q / W) % H), n = (int)(q / ((int64_t)W * H));
    float acc = 0.f;
    if (k == 0) {
      acc = dout[(int64_t)n * C + c] / (float)(H * W);
    } else {
      // windows oy with oy*s-p <= y < oy*s-p+k
      int oy0 = (y + p - k + s) / s; if (y + p - k + 1 <= 0) oy0 = 0;
      int ox0 = (x + p - k + s) / s; if (x + p - k + 1 <= 0) ox0 = 0;
      const int oy1 = min((y + p) / s, Ho - 1), ox1 = min((x + p) / s, Wo - 1);
      for (int oy = oy0; oy <= oy1; ++oy)
        for (int ox = ox0; ox <= ox1; ++ox)
          acc += dout[(((int64_t)n * Ho + oy) * Wo + ox) * C + c] /
                 (pool_inv(oy, k, s, p, H) * pool_inv(ox, k, s, p, W));
    }
    din[i] = rt(acc, rnd);
  }
}

// dst[p*dst_ld + dst_off + c] = src[p*src_ld + src_off + c]  (concat / slice along channels)
__global__ void __launch_bounds__(kT)
copy_channels_kernel(const float* __restrict__ src, int src_ld, int src_off, float* __restrict__ dst, int dst_ld,
                     int dst_off, int64_t npix, int C, int rnd) {
  const int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    dst[p * dst_ld + dst_off + c] = rt(src[p * src_ld + src_off + c], rnd);
  }
}

// torch.optim.SGD (momentum, weight decay, dampening 0, no nesterov) over one flat arena
__global__ void __launch_bounds__(kT)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n, float lr,
           float momentum, float wd, int first, float grad_scale) {
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT) {
    const float w = p[i];
    const float d = fmaf(wd, w, g[i] * grad_scale);
    const float b = first ? d : fmaf(momentum, buf[i], d);
    buf[i] = b;
    p[i] = w - lr * b;
  }
}

int conv_common(ConvArgs& a, const float* in, float* out, const float* w_packed, const float* bias, int N, int H,
                int W, int Cin, int Cout, int k, int stride) {
  a.in = in; a.in_dtype = LEDB200_F32; a.in_sc = 1; a.in_sw = Cin; a.in_sh = (int64_t)W * Cin;
  a.in_sn = (int64_t)H * W * Cin;
  a.out = out; a.out_dtype = LEDB200_F32; a.out_ld = Cout;
  a.bias = bias; a.w_direct = w_packed; a.cout_pad16 = (Cout + 15) / 16 * 16;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = k; a.stride = stride; a.pad = k / 2; a.dil = 1;
  a.Ho = (H + 2 * a.pad - k) / stride + 1; a.Wo = (W + 2 * a.pad - k) / stride + 1;
  return 0;
}

}  // namespace
}  // namespace ledb

using namespace ledb;

extern "C" {

int ledb200_train_set_tf32_rounding(int32_t on) {
  const int prev = g_round;
  g_round = on ? 1 : 0;
  return prev;
}

int ledb200_train_set_tf32_passes(int32_t passes) {
  if (passes != 1 && passes != 3) return fail(LEDB200_EINVAL, "train_set_tf32_passes: 1 or 3");
  const int prev = g_passes;
  g_passes = passes;
  return prev;
}

int64_t ledb200_peer_allreduce_buffer_bytes(int32_t nmax) {
  return nmax < 1 ? 0 : (int64_t)sizeof(double) * 2 * kPeerMax * nmax + (int64_t)sizeof(unsigned) * 2 * kPeerMax + 64;
}

int ledb200_peer_allreduce_f64(const double* local, int32_t n, int32_t rank, int32_t world, const uint64_t* peer_buffers,
                               uint32_t seq, int32_t nmax, double* out, void* stream) {
  if (!local || !out || !peer_buffers) return fail(LEDB200_EINVAL, "peer_allreduce: null buffer");
  if (world < 1 || world > kPeerMax || rank < 0 || rank >= world) return fail(LEDB200_EINVAL, "peer_allreduce: bad rank / world size");
  if (n < 1 || n > nmax) return fail(LEDB200_EINVAL, "peer_allreduce: vector longer than the peer slots");
  if (seq == 0) return fail(LEDB200_EINVAL, "peer_allreduce: sequence numbers start at 1 (the flags are zero-initialised)");
  PeerPtrs pp{};
  for (int i = 0; i < world; ++i) pp.p[i] = peer_buffers[i];
  peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(local, n, pp, rank, world, seq, nmax, out);
  LEDB_LAUNCH_OK("peer_allreduce_kernel");
  return LEDB200_OK;
}

// The stem's first layer on the NCHW image as the caller holds it (no NHWC copy): Conv2d(Cin <= 4 -> Cout % 16 == 0, 3x3,
// padding 1, stride 1 | 2).  y [N,Ho,Wo,Cout] NHWC; weights in the state dict's OIHW layout.
int ledb200_train_stem_fwd(const float* x_nchw, const float* w_oihw, const float* bias_opt, float* y, int32_t N, int32_t H,
                           int32_t W, int32_t Cin, int32_t Cout, int32_t stride, void* stream) {
  if (!x_nchw || !w_oihw || !y) return fail(LEDB200_EINVAL, "train_stem_fwd: null buffer");
  if (Cin < 1 || Cin > 4 || Cout < 16 || Cout % 16 || Cout > 128 || (stride != 1 && stride != 2))
    return fail(LEDB200_EINVAL, "train_stem_fwd: Cin <= 4, Cout a multiple of 16 (<= 128), stride 1 or 2");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const int64_t total = (int64_t)N * Ho * Wo;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(total, 128), 148 * 16));
  const size_t smem = sizeof(float) * 9 * Cin * Cout;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1) stem_fwd_kernel<1><<<grid, 128, smem, st>>>(x_nchw, w_oihw, bias_opt, y, N, H, W, Ho, Wo, Cout, stride);
  else if (Cin == 2) stem_fwd_kernel<2><<<grid, 128, smem, st>>>(x_nchw, w_oihw, bias_opt, y, N, H, W, Ho, Wo, Cout, stride);
  else if (Cin == 3) stem_fwd_kernel<3><<<grid, 128, smem, st>>>(x_nchw, w_oihw, bias_opt, y, N, H, W, Ho, Wo, Cout, stride);
  else stem_fwd_kernel<4><<<grid, 128, smem, st>>>(x_nchw, w_oihw, bias_opt, y, N, H, W, Ho, Wo, Cout, stride);
  LEDB_LAUNCH_OK("stem_fwd_kernel");
  return LEDB200_OK;
}

// its weight gradient, x again NCHW; workspace: ledb200_train_wgrad_workspace_bytes(Cin, Cout, 3)
int ledb200_train_stem_wgrad(const float* x_nchw, const float* dy, float* dw_oihw, int32_t N, int32_t H, int32_t W, int32_t Cin,
                             int32_t Cout, int32_t stride, void* workspace, void* stream) {
  if (!x_nchw || !dy || !dw_oihw || !workspace) return fail(LEDB200_EINVAL, "train_stem_wgrad: null buffer");
  if (Cin < 1 || Cin > 4 || Cout < 1 || (stride != 1 && stride != 2)) return fail(LEDB200_EINVAL, "train_stem_wgrad: Cin <= 4, stride 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const int gz = ceil_div(Cout, WG_CO);
  int64_t gx = (148 * 4) / gz;                     // slots the workspace was sized for: 2 * gx
  if (gx < 1) gx = 1;
  float* part = reinterpret_cast<float*>(reinterpret_cast<double*>(workspace) + bn_ws_doubles(Cout));
  const int64_t nw = (int64_t)Cout * Cin * 9;
  const int nseg = ceil_div(Wo, 64);
  const int64_t items = (int64_t)N * Ho * nseg;
  const int nslots = (int)std::max<int64_t>(1, std::min<int64_t>(2 * gx, std::min<int64_t>(ceil_div64(items, 8), 148 * 8)));
  dim3 g2((unsigned)nslots, (unsigned)ceil_div(Cout, 32));
  if (Cin == 1) wgrad_small_cin_kernel<1, true><<<g2, 256, 0, st>>>(x_nchw, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
  else if (Cin == 2) wgrad_small_cin_kernel<2, true><<<g2, 256, 0, st>>>(x_nchw, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
  else if (Cin == 3) wgrad_small_cin_kernel<3, true><<<g2, 256, 0, st>>>(x_nchw, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
  else wgrad_small_cin_kernel<4, true><<<g2, 256, 0, st>>>(x_nchw, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
  wgrad_sum_kernel<<<(unsigned)ceil_div64(nw, kT), kT, 0, st>>>(part, dw_oihw, nw, nslots);
  LEDB_LAUNCH_OK("train_stem_wgrad");
  return LEDB200_OK;
}

int ledb200_train_set_wgrad_passes(int32_t passes) {
  if (passes != 1 && passes != 3) return fail(LEDB200_EINVAL, "train_set_wgrad_passes: 1 or 3");
  const int prev = g_wgrad_passes;
  g_wgrad_passes = passes;
  return prev;
}

int64_t ledb200_train_packed_weight_floats(int32_t Cout, int32_t Cin, int32_t k, int32_t mode) {
  const int64_t taps = (int64_t)k * k;
  return mode == 0 ? taps * Cin * ((Cout + 15) / 16 * 16) : taps * Cout * ((Cin + 15) / 16 * 16);
}

int ledb200_train_pack_weight(const float* w_oihw, float* out, int32_t Cout, int32_t Cin, int32_t k, int32_t mode,
                              void* stream) {
  if (!w_oihw || !out) return fail(LEDB200_EINVAL, "pack_weight: null buffer");
  if (mode != 0 && mode != 1) return fail(LEDB200_EINVAL, "pack_weight: mode must be 0 (forward) or 1 (dgrad)");
  const int pad16 = mode == 0 ? (Cout + 15) / 16 * 16 : (Cin + 15) / 16 * 16;
  const int64_t total = ledb200_train_packed_weight_floats(Cout, Cin, k, mode);
  pack_weight_kernel<<<grid1d(total), kT, 0, (cudaStream_t)stream>>>(w_oihw, out, Cout, Cin, k * k, mode, pad16);
  LEDB_LAUNCH_OK("pack_weight_kernel");
  return LEDB200_OK;
}

int ledb200_train_conv_fwd(const float* x, const float* w_packed, const float* bias_opt, float* y, int32_t N,
                           int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* stream) {
  if (!x || !w_packed || !y) return fail(LEDB200_EINVAL, "train_conv_fwd: null buffer");
  if ((k != 1 && k != 3) || (stride != 1 && stride != 2)) return fail(LEDB200_EINVAL, "train_conv_fwd: k in {1,3}, stride in {1,2}");
  ConvArgs a;
  conv_common(a, x, y, w_packed, bias_opt, N, H, W, Cin, Cout, k, stride);
  return launch_conv_direct(a, (cudaStream_t)stream);
}

// dx [N,H,W,Cin] from dy [N,Ho,Wo,Cout]; w_packed_dgrad from ledb200_train_pack_weight(mode 1)
int ledb200_train_conv_dgrad(const float* dy, const float* w_packed_dgrad, float* dx, int32_t N, int32_t H, int32_t W,
                             int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* stream) {
  if (!dy || !w_packed_dgrad || !dx) return fail(LEDB200_EINVAL, "train_conv_dgrad: null buffer");
  if ((k != 1 && k != 3) || (stride != 1 && stride != 2)) return fail(LEDB200_EINVAL, "train_conv_dgrad: k in {1,3}, stride in {1,2}");
  const int pad = k / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  ConvArgs a;
  // a stride-1 convolution with Cin' = Cout, Cout' = Cin over the (zero-inserted) dY, virtual extent H x W
  a.in = dy; a.in_dtype = LEDB200_F32; a.in_sc = 1; a.in_sw = Cout; a.in_sh = (int64_t)Wo * Cout;
  a.in_sn = (int64_t)Ho * Wo * Cout;
  a.out = dx; a.out_dtype = LEDB200_F32; a.out_ld = Cin;
  a.w_direct = w_packed_dgrad; a.cout_pad16 = (Cin + 15) / 16 * 16;
  a.N = N; a.H = H; a.W = W; a.Cin = Cout; a.Cout = Cin; a.ksize = k; a.stride = 1; a.pad = pad; a.dil = 1;
  a.Ho = H; a.Wo = W;
  a.in_up = stride; a.Hr = Ho; a.Wr = Wo;
  if (stride == 1) { a.in_up = 1; }
  return launch_conv_direct(a, (cudaStream_t)stream);
}

// ---- tensor-core (TF32) forms of the two convolutions above: conv_tc.cu's implicit GEMM on fp32 NHWC tensors.
static void conv_tc_args(ConvArgs& a, const float* in, float* out, const float* w_tc, const float* bias, int N, int H, int W,
                         int Cin, int Cout, int k, int stride) {
  a.in = in; a.in_dtype = LEDB200_F32; a.in_sc = 1; a.in_sw = Cin; a.in_sh = (int64_t)W * Cin;
  a.in_sn = (int64_t)H * W * Cin;
  a.out = out; a.out_dtype = LEDB200_F32; a.out_ld = Cout;
  a.bias = bias; a.w_tc32 = w_tc; a.tf32 = g_passes; a.cout_pad_tc = conv_tc_pad(Cout);
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = k; a.stride = stride; a.pad = k / 2; a.dil = 1;
  a.Ho = (H + 2 * a.pad - k) / stride + 1; a.Wo = (W + 2 * a.pad - k) / stride + 1;
}

// 1 when the tensor-core kernels take this convolution (op 0 forward, 1 data gradient, 2 weight gradient), else 0:
// the caller then uses the CUDA-core entry points (odd sizes, Cin = 3, tiles that are not all interior).
int32_t ledb200_train_conv_tc_ok(int32_t op, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k,
                                 int32_t stride) {
  if (N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || (k != 1 && k != 3) || (stride != 1 && stride != 2)) return 0;
  ConvArgs a;
  const float* dummy = reinterpret_cast<const float*>(uintptr_t(256));
  if (op == 0) {
    conv_tc_args(a, dummy, const_cast<float*>(dummy), dummy, nullptr, N, H, W, Cin, Cout, k, stride);
    return conv_tc_eligible(a) ? 1 : 0;
  }
  if (op == 1) {
    if (stride == 2) {   // parity classes over dY's extents
      if ((H & 1) || (W & 1)) return 0;
      conv_tc_args(a, dummy, const_cast<float*>(dummy), dummy, nullptr, N, H / 2, W / 2, Cout, Cin, k, 1);
      a.sub = 1;
      return conv_tc_eligible(a) ? 1 : 0;
    }
    conv_tc_args(a, dummy, const_cast<float*>(dummy), dummy, nullptr, N, H, W, Cout, Cin, k, 1);
    return conv_tc_eligible(a) ? 1 : 0;
  }
  if (op == 2) return wgrad_tc_eligible(N, H, W, Cin, Cout, k, stride, g_wgrad_passes) ? 1 : 0;
  return 0;
}

int64_t ledb200_train_wgrad_tc_workspace_bytes(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k,
                                               int32_t stride) {
  return wgrad_tc_workspace_bytes(N, H, W, Cin, Cout, k, stride, g_wgrad_passes);
}

int ledb200_train_conv_wgrad_tc(const float* x, const float* dy, float* dw_oihw, int32_t N, int32_t H, int32_t W,
                                int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* workspace, void* stream) {
  if (!x || !dy || !dw_oihw || !workspace) return fail(LEDB200_EINVAL, "train_conv_wgrad_tc: null buffer");
  return launch_wgrad_tc(x, dy, dw_oihw, N, H, W, Cin, Cout, k, stride, g_wgrad_passes, workspace, (cudaStream_t)stream);
}

int64_t ledb200_train_packed_weight_tc_floats(int32_t Cout, int32_t Cin, int32_t k, int32_t mode) {
  const int64_t taps = (int64_t)k * k;
  return 2 * (mode == 0 ? (int64_t)conv_tc_pad(Cout) * taps * Cin : (int64_t)conv_tc_pad(Cin) * taps * Cout);   // modes 1, 2: same size
}

int ledb200_train_pack_weight_tc(const float* w_oihw, float* out, int32_t Cout, int32_t Cin, int32_t k, int32_t mode,
                                 void* stream) {
  if (!w_oihw || !out) return fail(LEDB200_EINVAL, "pack_weight_tc: null buffer");
  if (mode != 0 && mode != 1 && mode != 2)
    return fail(LEDB200_EINVAL, "pack_weight_tc: mode must be 0 (forward), 1 (dgrad) or 2 (dgrad of a 3x3 stride-2 convolution)");
  const int rows_pad = conv_tc_pad(mode == 0 ? Cout : Cin);
  if (mode == 2) {
    if (k != 3) return fail(LEDB200_EINVAL, "pack_weight_tc: mode 2 is for 3x3 filters");
    pack_weight_tc_sub_kernel<<<grid1d((int64_t)rows_pad * 9 * Cout), kT, 0, (cudaStream_t)stream>>>(w_oihw, out, Cout, Cin, rows_pad);
    LEDB_LAUNCH_OK("pack_weight_tc_sub_kernel");
    return LEDB200_OK;
  }
  const int64_t total = ledb200_train_packed_weight_tc_floats(Cout, Cin, k, mode) / 2;
  pack_weight_tc_kernel<<<grid1d(total), kT, 0, (cudaStream_t)stream>>>(w_oihw, out, Cout, Cin, k * k, mode, rows_pad);
  LEDB_LAUNCH_OK("pack_weight_tc_kernel");
  return LEDB200_OK;
}

int ledb200_train_conv_fwd_tc(const float* x, const float* w_tc, const float* bias_opt, float* y, int32_t N, int32_t H,
                              int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* stream) {
  if (!x || !w_tc || !y) return fail(LEDB200_EINVAL, "train_conv_fwd_tc: null buffer");
  if (!ledb200_train_conv_tc_ok(0, N, H, W, Cin, Cout, k, stride))
    return fail(LEDB200_EINVAL, "train_conv_fwd_tc: shape not eligible (ask ledb200_train_conv_tc_ok first)");
  ConvArgs a;
  conv_tc_args(a, x, y, w_tc, bias_opt, N, H, W, Cin, Cout, k, stride);
  return launch_conv_tc(a, (cudaStream_t)stream);
}

int ledb200_train_conv_dgrad_tc(const float* dy, const float* w_tc_dgrad, float* dx, int32_t N, int32_t H, int32_t W,
                                int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* stream) {
  if (!dy || !w_tc_dgrad || !dx) return fail(LEDB200_EINVAL, "train_conv_dgrad_tc: null buffer");
  if (!ledb200_train_conv_tc_ok(1, N, H, W, Cin, Cout, k, stride))
    return fail(LEDB200_EINVAL, "train_conv_dgrad_tc: shape not eligible (ask ledb200_train_conv_tc_ok first)");
  ConvArgs a;
  if (stride == 2) {
    // dX[2i + a, 2j + b] = sum over the filter rows / columns of that parity: four stride-1 convolutions over dY (3x3, weights
    // from pack mode 2), or one over the even-even class after zeroing dX (1x1, weights from pack mode 1)
    const int Ho = H / 2, Wo = W / 2;
    if (k == 1) {
      LEDB_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * H * W * Cin, (cudaStream_t)stream));
      conv_tc_args(a, dy, dx, w_tc_dgrad, nullptr, N, Ho, Wo, Cout, Cin, 1, 1);
      a.sub = 1;
      return launch_conv_tc(a, (cudaStream_t)stream);
    }
    const int before[4] = {0, 1, 3, 5};
    const int64_t rows_pad = conv_tc_pad(Cin);
    for (int cls = 0; cls < 4; ++cls) {
      conv_tc_args(a, dy, dx, w_tc_dgrad + 2 * rows_pad * Cout * before[cls], nullptr, N, Ho, Wo, Cout, Cin, 3, 1);
      a.sub = 1; a.sub_a = cls >> 1; a.sub_b = cls & 1;
      const int rc = launch_conv_tc(a, (cudaStream_t)stream);
      if (rc) return rc;
    }
    return LEDB200_OK;
  }
  // stride 1: dX = conv(dY, rotated weights with the channel roles swapped), same padding
  conv_tc_args(a, dy, dx, w_tc_dgrad, nullptr, N, H, W, Cout, Cin, k, 1);
  return launch_conv_tc(a, (cudaStream_t)stream);
}

// dw_oihw [Cout,Cin,k,k] (overwritten), dbias_opt [Cout] (overwritten); workspace >= 2*Cout doubles when dbias_opt
int64_t ledb200_train_bn_workspace_bytes(int32_t C) { return C < 1 ? 0 : 8 * bn_ws_doubles(C); }

int64_t ledb200_train_wgrad_workspace_bytes(int32_t Cin, int32_t Cout, int32_t k) {
  if (Cin < 1 || Cout < 1 || (k != 1 && k != 3)) return 0;
  const int gy = ceil_div(Cin, WG_CI), gz = ceil_div(Cout, WG_CO);
  int64_t gx = (148 * 4) / (gy * gz);
  if (gx < 1) gx = 1;
  return 8 * bn_ws_doubles(Cout) + 4 * 2 * gx * (int64_t)Cout * Cin * k * k;
}

int ledb200_train_conv_wgrad(const float* x, const float* dy, float* dw_oihw, float* dbias_opt, int32_t N, int32_t H,
                             int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* workspace,
                             void* stream) {
  if (!x || !dy || !dw_oihw) return fail(LEDB200_EINVAL, "train_conv_wgrad: null buffer");
  if ((k != 1 && k != 3) || (stride != 1 && stride != 2)) return fail(LEDB200_EINVAL, "train_conv_wgrad: k in {1,3}, stride in {1,2}");
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = k / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (!workspace) return fail(LEDB200_EINVAL, "train_conv_wgrad: workspace of ledb200_train_wgrad_workspace_bytes() needed");
  const int tiles_x = ceil_div(Wo, WG_TW), tiles_y = ceil_div(Ho, WG_TH);
  const int64_t total_tiles = (int64_t)N * tiles_x * tiles_y;
  const int gy = ceil_div(Cin, WG_CI), gz = ceil_div(Cout, WG_CO);
  int64_t gx = (148 * 4) / (gy * gz);
  if (gx < 1) gx = 1;
  if (gx > total_tiles) gx = total_tiles;
  const int IH = (WG_TH - 1) * stride + k, IW = (WG_TW - 1) * stride + k;
  const size_t smem = sizeof(float) * ((size_t)IH * IW * WG_CI + WG_TH * WG_TW * WG_CO);
  dim3 grid((unsigned)gx, gy, gz);
  // workspace: [bias-gradient accumulators and partials: bn_ws_doubles(Cout) doubles][2 * gx weight-gradient slots]
  float* part = reinterpret_cast<float*>(reinterpret_cast<double*>(workspace) + bn_ws_doubles(Cout));
  const int64_t nw = (int64_t)Cout * Cin * k * k;
  if (k == 3 && Cin <= 4 && (int64_t)N * Ho >= 64) {
    // narrow-input form (stem): one slot per block; slots fit the workspace sized for 2 * gx slots above
    const int nseg = ceil_div(Wo, 64);
    const int64_t items = (int64_t)N * Ho * nseg;
    const int nslots = (int)std::max<int64_t>(1, std::min<int64_t>(2 * gx, std::min<int64_t>(ceil_div64(items, 8), 148 * 8)));
    dim3 g2((unsigned)nslots, (unsigned)ceil_div(Cout, 32));
    if (Cin == 1) wgrad_small_cin_kernel<1><<<g2, 256, 0, st>>>(x, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
    else if (Cin == 2) wgrad_small_cin_kernel<2><<<g2, 256, 0, st>>>(x, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
    else if (Cin == 3) wgrad_small_cin_kernel<3><<<g2, 256, 0, st>>>(x, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
    else wgrad_small_cin_kernel<4><<<g2, 256, 0, st>>>(x, dy, part, N, H, W, Ho, Wo, Cout, stride, nseg);
    wgrad_sum_kernel<<<(unsigned)ceil_div64(nw, kT), kT, 0, st>>>(part, dw_oihw, nw, nslots);
    LEDB_LAUNCH_OK("wgrad_small_cin_kernel");
  } else
  if (k == 3) {
    LEDB_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_kernel<3><<<grid, 128, smem, st>>>(x, dy, part, N, H, W, Cin, Ho, Wo, Cout, stride, pad, tiles_x,
                                             tiles_x * tiles_y, total_tiles);
  } else {
    wgrad_kernel<1><<<grid, 128, smem, st>>>(x, dy, part, N, H, W, Cin, Ho, Wo, Cout, stride, pad, tiles_x,
                                             tiles_x * tiles_y, total_tiles);
  }
  if (!(k == 3 && Cin <= 4 && (int64_t)N * Ho >= 64))
    wgrad_sum_kernel<<<(unsigned)ceil_div64(nw, kT), kT, 0, st>>>(part, dw_oihw, nw, 2 * (int)gx);
  LEDB_LAUNCH_OK("wgrad_kernel");
  if (dbias_opt) {
    double* acc = (double*)workspace;
    const int64_t npix = (int64_t)N * Ho * Wo;
    int rc = chan_reduce<2>(dy, nullptr, nullptr, nullptr, nullptr, 0, npix, Cout, acc, st);
    if (rc) return rc;
    acc_to_float_kernel<<<ceil_div(Cout, 128), 128, 0, st>>>(acc, dbias_opt, Cout);
    LEDB_LAUNCH_OK("bias_grad");
  }
  return LEDB200_OK;
}

// ---- BatchNorm statistics as separate steps, so that a caller can all-reduce them between ranks (SyncBN,
//      configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:20: torch.nn.SyncBatchNorm semantics).
// mode 0: workspace[0:2C] = (sum y, sum y^2);  mode 1: workspace[0:2C] = (sum dz, sum dz * xhat), dz = masked dout.
int ledb200_train_bn_reduce(const float* a, const float* y_opt, const float* out_opt, const float* mean_opt,
                            const float* invstd_opt, int32_t mode, int32_t relu, int64_t npix, int32_t C,
                            void* workspace, void* stream) {
  if (!a || !workspace) return fail(LEDB200_EINVAL, "train_bn_reduce: null buffer");
  if (npix < 1 || C < 1) return fail(LEDB200_EINVAL, "train_bn_reduce: empty input");
  if ((size_t)C * 2 * sizeof(double) > 48 * 1024) return fail(LEDB200_EINVAL, "train_bn_reduce: C too large");
  if (mode != 0 && mode != 1 && mode != 3 && mode != 4) return fail(LEDB200_EINVAL, "train_bn_reduce: mode must be 0, 1, 3 or 4");
  if (mode != 0 && (!y_opt || !mean_opt || !invstd_opt || (relu && !out_opt)))
    return fail(LEDB200_EINVAL, "train_bn_reduce: backward statistics need y, mean, invstd (and out for the ReLU mask)");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  // mode 0 also leaves this rank's sample count at workspace[2C] (all-reduced together with the sums by a SyncBN caller);
  // mode 3 = mode 1 for a SyncBN caller that hands the FORWARD workspace back: the global count it still holds at [2C] moves
  // to [4C] (where bwd_apply reads it), and the sums are written twice, at [0, 2C) (rank-local) and [2C, 4C) (to be all-reduced);
  // mode 4 = mode 3 without the move (a second backward pass over the same graph: the count already sits at [4C])
  if (mode == 3) move_double_kernel<<<1, 1, 0, st>>>(acc + 4 * C, acc + 2 * C);
  int rc = mode == 0 ? chan_reduce<0>(a, nullptr, nullptr, nullptr, nullptr, 0, npix, C, acc, st, 0, 2 * C, (double)npix)
                     : chan_reduce<1>(a, y_opt, out_opt, mean_opt, invstd_opt, relu, npix, C, acc, st, mode >= 3 ? 2 * C : 0);
  if (rc) return rc;
  LEDB_LAUNCH_OK("train_bn_reduce");
  return LEDB200_OK;
}

// finalize + apply from statistics in workspace[0:2C] taken over `total_count` samples per channel (all ranks)
int ledb200_train_bn_fwd_apply(const float* y, const float* gamma, const float* beta, const float* res_opt, float* out,
                               float* save_mean, float* save_invstd, float* running_mean_opt, float* running_var_opt,
                               float momentum, float eps, int32_t relu, int64_t npix, double total_count, int32_t C,
                               const void* workspace, void* stream) {
  if (!y || !gamma || !beta || !out || !save_mean || !save_invstd || !workspace)
    return fail(LEDB200_EINVAL, "train_bn_fwd_apply: null buffer");
  if (npix < 1 || C < 1 || (total_count >= 0.0 && total_count < (double)npix))
    return fail(LEDB200_EINVAL, "train_bn_fwd_apply: bad sample count");
  cudaStream_t st = (cudaStream_t)stream;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>((const double*)workspace, C, total_count, eps, momentum, save_mean,
                                                       save_invstd, running_mean_opt, running_var_opt);
  if (vec4_ok(C, npix * C, y, res_opt, out, gamma, beta) && ((uintptr_t)save_mean | (uintptr_t)save_invstd) % 16 == 0)
    bn_apply4_kernel<<<grid1d(npix * C / 4), kT, 0, st>>>(
        reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(res_opt), gamma, beta, save_mean, save_invstd,
        reinterpret_cast<float4*>(out), relu, (uint32_t)(npix * C / 4), (uint32_t)(C / 4), g_round);
  else if ((npix * C) % 4 == 0 && npix * C < (1ll << 31) && C >= 4 && ((uintptr_t)y | (uintptr_t)res_opt | (uintptr_t)out) % 16 == 0)
    bn_apply_flat4_kernel<<<grid1d(npix * C / 4), kT, 0, st>>>(
        reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(res_opt), gamma, beta, save_mean, save_invstd,
        reinterpret_cast<float4*>(out), relu, (uint32_t)(npix * C / 4), (uint32_t)C, g_round);
  else if (npix * C < (1ll << 32))
    bn_apply_kernel<uint32_t><<<grid1d(npix * C), kT, 0, st>>>(y, res_opt, gamma, beta, save_mean, save_invstd, out, relu,
                                                               npix * C, C, g_round);
  else
    bn_apply_kernel<int64_t><<<grid1d(npix * C), kT, 0, st>>>(y, res_opt, gamma, beta, save_mean, save_invstd, out, relu,
                                                              npix * C, C, g_round);
  LEDB_LAUNCH_OK("train_bn_fwd_apply");
  return LEDB200_OK;
}

// backward from statistics: workspace[0:2C] = this rank's (sum dz, sum dz*xhat) -> dgamma / dbeta,
// workspace[2C:4C] = the same summed over all ranks -> dy (1 / total_count)
int ledb200_train_bn_bwd_apply(const float* dout, const float* y, const float* out, const float* gamma,
                               const float* save_mean, const float* save_invstd, float* dy, float* dres_opt,
                               float* dgamma, float* dbeta, int32_t relu, int64_t npix, double total_count, int32_t C,
                               const void* workspace, void* stream) {
  if (!dout || !y || !gamma || !save_mean || !save_invstd || !dy || !dgamma || !dbeta || !workspace)
    return fail(LEDB200_EINVAL, "train_bn_bwd_apply: null buffer");
  if (relu && !out) return fail(LEDB200_EINVAL, "train_bn_bwd_apply: the ReLU mask needs the forward output");
  if (total_count >= 0.0 && total_count < (double)npix) return fail(LEDB200_EINVAL, "train_bn_bwd_apply: bad sample count");
  const double* acc = (const double*)workspace;
  if (vec4_ok(C, npix * C, dout, y, out, dy, dres_opt) && ((uintptr_t)gamma | (uintptr_t)save_mean | (uintptr_t)save_invstd) % 16 == 0)
    bn_bwd_apply4_kernel<<<grid1d(npix * C / 4), kT, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(out), gamma,
        save_mean, save_invstd, acc + 2 * C, acc, reinterpret_cast<float4*>(dy), reinterpret_cast<float4*>(dres_opt), dgamma, dbeta,
        relu, (uint32_t)(npix * C / 4), (uint32_t)(C / 4), total_count < 0.0 ? -1.f : (float)(1.0 / total_count), g_round);
  else if ((npix * C) % 4 == 0 && npix * C < (1ll << 31) && C >= 4 &&
           ((uintptr_t)dout | (uintptr_t)y | (uintptr_t)out | (uintptr_t)dy | (uintptr_t)dres_opt) % 16 == 0)
    bn_bwd_apply_flat4_kernel<<<grid1d(npix * C / 4), kT, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(out), gamma,
        save_mean, save_invstd, acc + 2 * C, acc, reinterpret_cast<float4*>(dy), reinterpret_cast<float4*>(dres_opt), dgamma, dbeta,
        relu, (uint32_t)(npix * C / 4), (uint32_t)C, total_count < 0.0 ? -1.f : (float)(1.0 / total_count), g_round);
  else if (npix * C < (1ll << 32))
    bn_bwd_apply_kernel<uint32_t><<<grid1d(npix * C), kT, 0, (cudaStream_t)stream>>>(
        dout, y, out, gamma, save_mean, save_invstd, acc + 2 * C, acc, dy, dres_opt, dgamma, dbeta, relu, npix * C, C,
        total_count < 0.0 ? -1.f : (float)(1.0 / total_count), g_round);
  else
    bn_bwd_apply_kernel<int64_t><<<grid1d(npix * C), kT, 0, (cudaStream_t)stream>>>(
        dout, y, out, gamma, save_mean, save_invstd, acc + 2 * C, acc, dy, dres_opt, dgamma, dbeta, relu, npix * C, C,
        total_count < 0.0 ? -1.f : (float)(1.0 / total_count), g_round);
  LEDB_LAUNCH_OK("train_bn_bwd_apply");
  return LEDB200_OK;
}

// BatchNorm2d in training mode (+ residual add, + ReLU): out = [relu](bn(y) [+ res]).
// save_mean/save_invstd [C] are outputs (needed by backward); running stats updated in place when given.
// workspace: ledb200_train_bn_workspace_bytes(C).
int ledb200_train_bn_fwd(const float* y, const float* gamma, const float* beta, const float* res_opt, float* out,
                         float* save_mean, float* save_invstd, float* running_mean_opt, float* running_var_opt,
                         float momentum, float eps, int32_t relu, int64_t npix, int32_t C, void* workspace,
                         void* stream) {
  if (!y || !gamma || !beta || !out || !save_mean || !save_invstd || !workspace)
    return fail(LEDB200_EINVAL, "train_bn_fwd: null buffer");
  int rc = ledb200_train_bn_reduce(y, nullptr, nullptr, nullptr, nullptr, 0, 0, npix, C, workspace, stream);
  if (rc) return rc;
  return ledb200_train_bn_fwd_apply(y, gamma, beta, res_opt, out, save_mean, save_invstd, running_mean_opt,
                                    running_var_opt, momentum, eps, relu, npix, (double)npix, C, workspace, stream);
}

int ledb200_train_bn_bwd(const float* dout, const float* y, const float* out, const float* gamma,
                         const float* save_mean, const float* save_invstd, float* dy, float* dres_opt, float* dgamma,
                         float* dbeta, int32_t relu, int64_t npix, int32_t C, void* workspace, void* stream) {
  if (!dout || !y || !gamma || !save_mean || !save_invstd || !dy || !dgamma || !dbeta || !workspace)
    return fail(LEDB200_EINVAL, "train_bn_bwd: null buffer");
  if (relu && !out) return fail(LEDB200_EINVAL, "train_bn_bwd: the ReLU mask needs the forward output");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  int rc = chan_reduce<1>(dout, y, out, save_mean, save_invstd, relu, npix, C, acc, st);
  if (rc) return rc;
  if (vec4_ok(C, npix * C, dout, y, out, dy, dres_opt) && ((uintptr_t)gamma | (uintptr_t)save_mean | (uintptr_t)save_invstd) % 16 == 0)
    bn_bwd_apply4_kernel<<<grid1d(npix * C / 4), kT, 0, st>>>(
        reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(out), gamma,
        save_mean, save_invstd, acc, acc, reinterpret_cast<float4*>(dy), reinterpret_cast<float4*>(dres_opt), dgamma, dbeta, relu,
        (uint32_t)(npix * C / 4), (uint32_t)(C / 4), 1.f / (float)npix, g_round);
  else if ((npix * C) % 4 == 0 && npix * C < (1ll << 31) && C >= 4 &&
           ((uintptr_t)dout | (uintptr_t)y | (uintptr_t)out | (uintptr_t)dy | (uintptr_t)dres_opt) % 16 == 0)
    bn_bwd_apply_flat4_kernel<<<grid1d(npix * C / 4), kT, 0, st>>>(
        reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(out), gamma,
        save_mean, save_invstd, acc, acc, reinterpret_cast<float4*>(dy), reinterpret_cast<float4*>(dres_opt), dgamma, dbeta, relu,
        (uint32_t)(npix * C / 4), (uint32_t)C, 1.f / (float)npix, g_round);
  else if (npix * C < (1ll << 32))
    bn_bwd_apply_kernel<uint32_t><<<grid1d(npix * C), kT, 0, st>>>(dout, y, out, gamma, save_mean, save_invstd, acc, acc, dy,
                                                                   dres_opt, dgamma, dbeta, relu, npix * C, C,
                                                                   1.f / (float)npix, g_round);
  else
    bn_bwd_apply_kernel<int64_t><<<grid1d(npix * C), kT, 0, st>>>(dout, y, out, gamma, save_mean, save_invstd, acc, acc, dy,
                                                                  dres_opt, dgamma, dbeta, relu, npix * C, C,
                                                                  1.f / (float)npix, g_round);
  LEDB_LAUNCH_OK("train_bn_bwd");
  return LEDB200_OK;
}

int ledb200_train_resize_fwd(const float* src, float* out, int32_t N, int32_t h, int32_t w, int32_t H, int32_t W,
                             int32_t C, void* stream) {
  if (!src || !out) return fail(LEDB200_EINVAL, "train_resize_fwd: null buffer");
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  if ((int64_t)N * H < (1ll << 31) && (int64_t)std::max(W, w) * C < (1ll << 30)) {
    const int V = (C % 4 == 0 && ((uintptr_t)src | (uintptr_t)out) % 16 == 0) ? 4 : 1;
    dim3 grid((unsigned)(N * H), (unsigned)std::min(8, ceil_div(W * C, kT * V * 4)));
    if (V == 4) resize_fwd_row_kernel<4><<<grid, kT, 0, (cudaStream_t)stream>>>(src, out, h, w, H, W, C, sh, sw, g_round);
    else resize_fwd_row_kernel<1><<<grid, kT, 0, (cudaStream_t)stream>>>(src, out, h, w, H, W, C, sh, sw, g_round);
  } else
  resize_fwd_kernel<<<grid1d((int64_t)N * H * W * C), kT, 0, (cudaStream_t)stream>>>(
      src, out, N, h, w, H, W, C, sh, sw, g_round);
  LEDB_LAUNCH_OK("resize_fwd_kernel");
  return LEDB200_OK;
}

int ledb200_train_resize_bwd(const float* dout, float* dsrc, int32_t N, int32_t h, int32_t w, int32_t H, int32_t W,
                             int32_t C, void* stream) {
  if (!dout || !dsrc) return fail(LEDB200_EINVAL, "train_resize_bwd: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  if ((int64_t)N * h < (1ll << 31) && (int64_t)std::max(W, w) * C < (1ll << 30) && (size_t)w * sizeof(BwdCol) <= 40 * 1024) {
    const int V = (C % 4 == 0 && ((uintptr_t)dout | (uintptr_t)dsrc) % 16 == 0) ? 4 : 1;
    // one block per source row when the row is short (the column table is per block), up to 4 otherwise
    dim3 grid((unsigned)(N * h), (unsigned)std::min(4, ceil_div(w * C, kT * V * 8)));
    const size_t smem = (size_t)w * sizeof(BwdCol);
    if (V == 4) resize_bwd_row_kernel<4><<<grid, kT, smem, st>>>(dout, dsrc, h, w, H, W, C, sh, sw, g_round);
    else resize_bwd_row_kernel<1><<<grid, kT, smem, st>>>(dout, dsrc, h, w, H, W, C, sh, sw, g_round);
  } else
  resize_bwd_kernel<<<grid1d((int64_t)N * h * w * C), kT, 0, st>>>(dout, dsrc, N, h, w, H, W, C, sh, sw, g_round);
  LEDB_LAUNCH_OK("resize_bwd_kernel");
  return LEDB200_OK;
}

int ledb200_train_add_relu(const float* a, const float* b_opt, float* out, int32_t relu, int64_t n, void* stream) {
  if (!a || !out) return fail(LEDB200_EINVAL, "train_add_relu: null buffer");
  add_relu_kernel<<<grid1d(n), kT, 0, (cudaStream_t)stream>>>(a, b_opt, out, relu, n, g_round);
  LEDB_LAUNCH_OK("add_relu_kernel");
  return LEDB200_OK;
}

int ledb200_train_relu_bwd(const float* dout, const float* out, float* dx, int64_t n, void* stream) {
  if (!dout || !out || !dx) return fail(LEDB200_EINVAL, "train_relu_bwd: null buffer");
  relu_bwd_kernel<<<grid1d(n), kT, 0, (cudaStream_t)stream>>>(dout, out, dx, n, g_round);
  LEDB_LAUNCH_OK("relu_bwd_kernel");
  return LEDB200_OK;
}

int ledb200_train_avgpool_fwd(const float* in, float* out, int32_t N, int32_t H, int32_t W, int32_t C, int32_t Ho,
                              int32_t Wo, int32_t k, int32_t s, int32_t p, void* stream) {
  if (!in || !out) return fail(LEDB200_EINVAL, "train_avgpool_fwd: null buffer");
  avgpool_fwd_kernel<<<grid1d((int64_t)N * Ho * Wo * C), kT, 0, (cudaStream_t)stream>>>(in, out, N, H, W, C, Ho, Wo, k,
                                                                                      s, p, g_round);
  LEDB_LAUNCH_OK("avgpool_fwd_kernel");
  return LEDB200_OK;
}

int ledb200_train_avgpool_bwd(const float* dout, float* din, int32_t N, int32_t H, int32_t W, int32_t C, int32_t Ho,
                              int32_t Wo, int32_t k, int32_t s, int32_t p, void* stream) {
  if (!dout || !din) return fail(LEDB200_EINVAL, "train_avgpool_bwd: null buffer");
  avgpool_bwd_kernel<<<grid1d((int64_t)N * H * W * C), kT, 0, (cudaStream_t)stream>>>(dout, din, N, H, W, C, Ho, Wo, k,
                                                                                    s, p, g_round);
  LEDB_LAUNCH_OK("avgpool_bwd_kernel");
  return LEDB200_OK;
}

int ledb200_train_copy_channels(const float* src, int32_t src_ld, int32_t src_off, float* dst, int32_t dst_ld,
                                int32_t dst_off, int64_t npix, int32_t C, void* stream) {
  if (!src || !dst) return fail(LEDB200_EINVAL, "train_copy_channels: null buffer");
  if (src_off + C > src_ld || dst_off + C > dst_ld) return fail(LEDB200_EINVAL, "train_copy_channels: slice out of range");
  copy_channels_kernel<<<grid1d(npix * C), kT, 0, (cudaStream_t)stream>>>(src, src_ld, src_off, dst, dst_ld, dst_off,
                                                                        npix, C, g_round);
  LEDB_LAUNCH_OK("copy_channels_kernel");
  return LEDB200_OK;
}

int ledb200_train_layout(const float* in, float* out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t to_nhwc,
                         void* stream) {
  if (!in || !out) return fail(LEDB200_EINVAL, "train_layout: null buffer");
  if (to_nhwc) return launch_nchw_to_nhwc(in, out, LEDB200_F32, N, C, H, W, (cudaStream_t)stream);
  return launch_nhwc_to_nchw(in, LEDB200_F32, out, N, C, H, W, C, (cudaStream_t)stream);
}

int ledb200_train_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                           float weight_decay, int32_t first_step, float grad_scale, void* stream) {
  if (!param || !grad || !momentum_buf) return fail(LEDB200_EINVAL, "train_sgd_step: null buffer");
  sgd_kernel<<<grid1d(n), kT, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, n, lr, momentum, weight_decay,
                                                         first_step, grad_scale);
  LEDB_LAUNCH_OK("sgd_kernel");
  return LEDB200_OK;
}

}  // extern "C"
