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
#include "tc_common.cuh"

namespace ledb {
namespace {

constexpr int kT = 256;
// tf32 storage mode (ledb200_train_set_tf32_rounding): every activation / gradient tensor these kernels write is rounded to
// tf32 (round to nearest, cvt.rna) so that a tensor-core consumer, which TRUNCATES raw fp32 operands to tf32 (measured:
// tests/test_gpu_train_tc.py, -3.7e-4 mean), sees exactly representable values - truncation would shrink every data
// gradient by ~3e-4 per layer, compounding to percents at the stem.  Off: plain fp32 stores (the CUDA-core path).
int g_round = 0;
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

// K-major fp32 weight matrix of the tensor-core path (conv_tc.cu, TF32 instantiations), values rounded to tf32 (RN):
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
    out[i] = __uint_as_float(t);
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
        s0 += dz; s1 = fmaf(dz, (y[i] - m) * is, s1);
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

__global__ void __launch_bounds__(kT) chan_sum_kernel(const double* __restrict__ part, int nblocks, int n2c,
                                                      double* __restrict__ acc) {
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= n2c) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * n2c + i];
  acc[i] = s;
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, double npix, float eps, float momentum,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
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
__global__ void __launch_bounds__(kT)
bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ res, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                float* __restrict__ out, int relu, int64_t total, int C, int rnd) {
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    float v = fmaf((y[i] - mean[c]) * invstd[c], gamma[c], beta[c]);
    if (res) v += res[i];
    out[i] = rt(relu ? fmaxf(v, 0.f) : v, rnd);
  }
}

// dy = gamma * invstd * (dz - dbeta/M - xhat * dgamma/M); dres = dz
__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ out,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const double* __restrict__ acc, const double* __restrict__ acc_local, float* __restrict__ dy,
                    float* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int relu, int64_t total,
                    int C, float invM, int rnd) {
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % C);
    float dz = dout[i];
    if (relu && !(out[i] > 0.f)) dz = 0.f;
    const float db = (float)acc[c], dg = (float)acc[C + c];
    const float is = invstd[c];
    const float xhat = (y[i] - mean[c]) * is;
    dy[i] = rt(gamma[c] * is * (dz - db * invM - xhat * dg * invM), rnd);
    if (dres) dres[i] = dz;
    if (i < C) { dbeta[c] = (float)acc_local[c]; dgamma[c] = (float)acc_local[C + c]; }   // parameter gradients stay rank-local (DDP averages them)
  }
}

__global__ void acc_to_float_kernel(const double* __restrict__ acc, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[c];
}

// workspace layout of the per-channel reductions (doubles): [0, 4C + 8) accumulators (2C of a reduction; SyncBN keeps a second
// copy and the sample count there), then kMaxRedBlocks block partials of 2C each
inline int64_t bn_ws_doubles(int C) { return 4 * (int64_t)C + 8 + (int64_t)kMaxRedBlocks * 2 * C; }

template <int KIND>
int chan_reduce(const float* a, const float* y, const float* out, const float* mean, const float* invstd, int relu,
                int64_t npix, int C, double* ws, cudaStream_t st) {
  double* part = ws + 4 * (int64_t)C + 8;
  const int grid = grid1d(npix * C, 4);                     // <= kMaxRedBlocks
  chan_reduce_kernel<KIND><<<grid, kT, sizeof(double) * 2 * C, st>>>(a, y, out, mean, invstd, relu, npix, C, part);
  chan_sum_kernel<<<ceil_div(2 * C, kT), kT, 0, st>>>(part, grid, 2 * C, ws);
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
  a.bias = bias; a.w_tc32 = w_tc; a.tf32 = 1; a.cout_pad_tc = conv_tc_pad(Cout);
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
    if (stride != 1) return 0;
    conv_tc_args(a, dummy, const_cast<float*>(dummy), dummy, nullptr, N, H, W, Cout, Cin, k, 1);
    return conv_tc_eligible(a) ? 1 : 0;
  }
  if (op == 2) return wgrad_tc_eligible(N, H, W, Cin, Cout, k, stride) ? 1 : 0;
  return 0;
}

int64_t ledb200_train_wgrad_tc_workspace_bytes(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k,
                                               int32_t stride) {
  return wgrad_tc_workspace_bytes(N, H, W, Cin, Cout, k, stride);
}

int ledb200_train_conv_wgrad_tc(const float* x, const float* dy, float* dw_oihw, int32_t N, int32_t H, int32_t W,
                                int32_t Cin, int32_t Cout, int32_t k, int32_t stride, void* workspace, void* stream) {
  if (!x || !dy || !dw_oihw || !workspace) return fail(LEDB200_EINVAL, "train_conv_wgrad_tc: null buffer");
  return launch_wgrad_tc(x, dy, dw_oihw, N, H, W, Cin, Cout, k, stride, workspace, (cudaStream_t)stream);
}

int64_t ledb200_train_packed_weight_tc_floats(int32_t Cout, int32_t Cin, int32_t k, int32_t mode) {
  const int64_t taps = (int64_t)k * k;
  return mode == 0 ? (int64_t)conv_tc_pad(Cout) * taps * Cin : (int64_t)conv_tc_pad(Cin) * taps * Cout;
}

int ledb200_train_pack_weight_tc(const float* w_oihw, float* out, int32_t Cout, int32_t Cin, int32_t k, int32_t mode,
                                 void* stream) {
  if (!w_oihw || !out) return fail(LEDB200_EINVAL, "pack_weight_tc: null buffer");
  if (mode != 0 && mode != 1) return fail(LEDB200_EINVAL, "pack_weight_tc: mode must be 0 (forward) or 1 (dgrad)");
  const int rows_pad = conv_tc_pad(mode == 0 ? Cout : Cin);
  const int64_t total = ledb200_train_packed_weight_tc_floats(Cout, Cin, k, mode);
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
  ConvArgs a;   // stride 1: dX = conv(dY, rotated weights with the channel roles swapped), same padding
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
  if (k == 3) {
    LEDB_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_kernel<3><<<grid, 128, smem, st>>>(x, dy, part, N, H, W, Cin, Ho, Wo, Cout, stride, pad, tiles_x,
                                             tiles_x * tiles_y, total_tiles);
  } else {
    wgrad_kernel<1><<<grid, 128, smem, st>>>(x, dy, part, N, H, W, Cin, Ho, Wo, Cout, stride, pad, tiles_x,
                                             tiles_x * tiles_y, total_tiles);
  }
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
  if (mode == 1 && (!y_opt || !mean_opt || !invstd_opt || (relu && !out_opt)))
    return fail(LEDB200_EINVAL, "train_bn_reduce: backward statistics need y, mean, invstd (and out for the ReLU mask)");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  int rc = mode == 0 ? chan_reduce<0>(a, nullptr, nullptr, nullptr, nullptr, 0, npix, C, acc, st)
                     : chan_reduce<1>(a, y_opt, out_opt, mean_opt, invstd_opt, relu, npix, C, acc, st);
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
  if (npix < 1 || C < 1 || total_count < (double)npix) return fail(LEDB200_EINVAL, "train_bn_fwd_apply: bad sample count");
  cudaStream_t st = (cudaStream_t)stream;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>((const double*)workspace, C, total_count, eps, momentum, save_mean,
                                                       save_invstd, running_mean_opt, running_var_opt);
  bn_apply_kernel<<<grid1d(npix * C), kT, 0, st>>>(y, res_opt, gamma, beta, save_mean, save_invstd, out, relu,
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
  if (total_count < (double)npix) return fail(LEDB200_EINVAL, "train_bn_bwd_apply: bad sample count");
  const double* acc = (const double*)workspace;
  bn_bwd_apply_kernel<<<grid1d(npix * C), kT, 0, (cudaStream_t)stream>>>(dout, y, out, gamma, save_mean, save_invstd,
                                                                         acc + 2 * C, acc, dy, dres_opt, dgamma, dbeta,
                                                                         relu, npix * C, C, (float)(1.0 / total_count), g_round);
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
  bn_bwd_apply_kernel<<<grid1d(npix * C), kT, 0, st>>>(dout, y, out, gamma, save_mean, save_invstd, acc, acc, dy, dres_opt,
                                                       dgamma, dbeta, relu, npix * C, C, 1.f / (float)npix, g_round);
  LEDB_LAUNCH_OK("train_bn_bwd");
  return LEDB200_OK;
}

int ledb200_train_resize_fwd(const float* src, float* out, int32_t N, int32_t h, int32_t w, int32_t H, int32_t W,
                             int32_t C, void* stream) {
  if (!src || !out) return fail(LEDB200_EINVAL, "train_resize_fwd: null buffer");
  resize_fwd_kernel<<<grid1d((int64_t)N * H * W * C), kT, 0, (cudaStream_t)stream>>>(
      src, out, N, h, w, H, W, C, (float)h / (float)H, (float)w / (float)W, g_round);
  LEDB_LAUNCH_OK("resize_fwd_kernel");
  return LEDB200_OK;
}

int ledb200_train_resize_bwd(const float* dout, float* dsrc, int32_t N, int32_t h, int32_t w, int32_t H, int32_t W,
                             int32_t C, void* stream) {
  if (!dout || !dsrc) return fail(LEDB200_EINVAL, "train_resize_bwd: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  resize_bwd_kernel<<<grid1d((int64_t)N * h * w * C), kT, 0, st>>>(dout, dsrc, N, h, w, H, W, C, (float)h / (float)H,
                                                                  (float)w / (float)W, g_round);
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
