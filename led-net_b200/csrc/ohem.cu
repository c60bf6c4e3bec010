// OHEM cross-entropy forward + backward and top-1 accuracy (north_star kernel 6, loss part).
//
// Replaces OhemCrossEntropy.forward (mmseg/models/losses/ohem_cross_entropy_loss.py:52-90):
//   softmax -> per-pixel (class-weighted) CE with ignore -> probability of the true class ->
//   FULL ascending sort of N*H*W floats only to read sorted[min(min_kept, n-1)] ->
//   threshold = max(that, thres) -> plain mean of CE over valid pixels with prob < threshold,
// its autograd backward, and accuracy() (mmseg/models/losses/accuracy.py:41-60) on the same logits.
// The sort is replaced by an exact 3-pass radix select over the float bit patterns
// (probabilities are >= 0, so their uint32 bit patterns are ordered like the values); the
// comparison `prob < threshold` is strict, so ties and order do not matter and selection is exact.
// The masked mean uses a fixed grid and fixed-order block partials (deterministic).
//
// logits fp32 NCHW (what LEDHead.loss_by_feat hands the loss), target int64.
#include <algorithm>
#include <cmath>

#include "kernels.h"

namespace ledb {
namespace {

constexpr int kT = 256;
constexpr int kPartials = 148 * 4;   // fixed number of reduction blocks
constexpr int kBins = 2048;

struct OhemState {
  unsigned long long nvalid;
  unsigned long long ncorrect;
  long long k;               // remaining rank inside the current prefix bucket
  unsigned int prefix;       // selected high bits so far
  unsigned int mask;         // which bits of `prefix` are fixed
  float threshold;
  float kept;
  unsigned int hist[3][kBins];
  double psum[kPartials];
  unsigned long long pcnt[kPartials];
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < kT / 32) ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;   // valid in warp 0
}

__global__ void ohem_init_kernel(OhemState* s) {
  for (int i = threadIdx.x; i < 3 * kBins; i += blockDim.x) (&s->hist[0][0])[i] = 0;
  if (threadIdx.x == 0) {
    s->nvalid = 0; s->ncorrect = 0; s->k = 0; s->prefix = 0; s->mask = 0; s->threshold = 0.f; s->kept = 0.f;
  }
}

// per pixel: softmax prob of the true class, weighted NLL, top-1 correctness
__global__ void __launch_bounds__(kT)
ohem_pixel_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int K, int64_t HW,
                  int64_t npix, int ignore, const float* __restrict__ cw, float* __restrict__ prob,
                  float* __restrict__ loss, OhemState* s) {
  __shared__ float sh[kT / 32];
  float nvalid = 0.f, ncorrect = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < npix; i += (int64_t)gridDim.x * kT) {
    const int64_t n = i / HW, hw = i % HW;
    const float* x = logits + n * K * HW + hw;
    const int64_t y = target[i];
    float m = -INFINITY;
    int am = 0;
    for (int k = 0; k < K; ++k) {
      const float v = __ldg(x + k * HW);
      if (v > m) { m = v; am = k; }
    }
    float p = 2.0f, l = 0.f;          // 2.0 = sentinel for ignored pixels (sorts after every prob)
    if (y != ignore) {
      float sum = 0.f;
      for (int k = 0; k < K; ++k) sum += expf(__ldg(x + k * HW) - m);
      const int yy = (y >= 0 && y < K) ? (int)y : 0;
      const float xy = __ldg(x + yy * HW) - m;
      p = expf(xy) / sum;
      l = -(xy - logf(sum)) * (cw ? cw[yy] : 1.f);
      nvalid += 1.f;
      ncorrect += (am == (int)y) ? 1.f : 0.f;
    }
    prob[i] = p;
    loss[i] = l;
  }
  const float a = block_sum(nvalid, sh);
  const float b = block_sum(ncorrect, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&s->nvalid, (unsigned long long)a);
    atomicAdd(&s->ncorrect, (unsigned long long)b);
  }
}

// rank k = min(min_kept, nvalid-1)
__global__ void ohem_rank_kernel(OhemState* s, long long min_kept) {
  const long long nv = (long long)s->nvalid;
  s->k = nv > 0 ? (min_kept < nv - 1 ? min_kept : nv - 1) : 0;
}

__global__ void __launch_bounds__(kT)
ohem_hist_kernel(const float* __restrict__ prob, int64_t npix, OhemState* s, int pass, int shift, int nbins) {
  __shared__ unsigned int h[kBins];
  for (int i = threadIdx.x; i < nbins; i += kT) h[i] = 0;
  __syncthreads();
  const unsigned prefix = s->prefix, mask = s->mask;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < npix; i += (int64_t)gridDim.x * kT) {
    const unsigned bits = __float_as_uint(prob[i]);
    if ((bits & mask) == prefix) atomicAdd(&h[(bits >> shift) & (nbins - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += kT)
    if (h[i]) atomicAdd(&s->hist[pass][i], h[i]);
}

// single thread: walk the histogram to the bucket holding rank k
__global__ void ohem_find_kernel(OhemState* s, int pass, int shift, int nbins, float thres) {
  long long k = s->k, cum = 0;
  int b = 0;
  for (; b < nbins; ++b) {
    const long long c = s->hist[pass][b];
    if (k < cum + c) break;
    cum += c;
  }
  if (b == nbins) b = nbins - 1;
  s->k = k - cum;
  s->prefix |= ((unsigned)b) << shift;
  s->mask |= ((unsigned)(nbins - 1)) << shift;
  if (pass == 2) {
    const float kth = __uint_as_float(s->prefix);
    s->threshold = fmaxf(kth, thres);   // ohem_cross_entropy_loss.py:86
  }
}

__global__ void __launch_bounds__(kT)
ohem_reduce_kernel(const float* __restrict__ prob, const float* __restrict__ loss, int64_t npix, OhemState* s) {
  __shared__ float sh[kT / 32];
  const float thr = s->threshold;
  // contiguous slab per block -> fixed summation order
  const int64_t per = ceil_div64(npix, kPartials);
  const int64_t lo = blockIdx.x * per, hi = (lo + per < npix) ? lo + per : npix;
  float sum = 0.f, cnt = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kT) {
    const float p = prob[i];
    if (p < thr && p <= 1.5f) { sum += loss[i]; cnt += 1.f; }
  }
  const float a = block_sum(sum, sh);
  const float b = block_sum(cnt, sh);
  if (threadIdx.x == 0) { s->psum[blockIdx.x] = (double)a; s->pcnt[blockIdx.x] = (unsigned long long)b; }
}

__global__ void ohem_final_kernel(OhemState* s, float loss_weight, float* out3) {
  double sum = 0.0;
  unsigned long long cnt = 0;
  for (int i = 0; i < kPartials; ++i) { sum += s->psum[i]; cnt += s->pcnt[i]; }
  const float eps = 1.1920929e-07f;   // torch.finfo(float32).eps (accuracy.py:52)
  s->kept = (float)cnt;
  if (s->nvalid == 0) {
    out3[0] = 0.f;                     // ohem_cross_entropy_loss.py:83-84
  } else {
    out3[0] = loss_weight * (float)(sum / (double)cnt);   // 0/0 -> NaN like .mean() of an empty tensor
  }
  out3[1] = (float)cnt;
  out3[2] = ((float)s->ncorrect + eps) * (100.0f / ((float)s->nvalid + eps));
}

__global__ void __launch_bounds__(kT)
ohem_backward_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                     const float* __restrict__ prob, int K, int64_t HW, int64_t npix, int ignore,
                     const float* __restrict__ cw, float loss_weight, const OhemState* __restrict__ s,
                     float* __restrict__ dlogits) {
  const float thr = s->threshold;
  const float scale = loss_weight / s->kept;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < npix; i += (int64_t)gridDim.x * kT) {
    const int64_t n = i / HW, hw = i % HW;
    const float* x = logits + n * K * HW + hw;
    float* d = dlogits + n * K * HW + hw;
    const float p = prob[i];
    const bool keep = (p < thr) && (p <= 1.5f);
    if (!keep) {
      for (int k = 0; k < K; ++k) d[k * HW] = 0.f;
      continue;
    }
    const int64_t y = target[i];
    const int yy = (y >= 0 && y < K) ? (int)y : 0;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, __ldg(x + k * HW));
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += expf(__ldg(x + k * HW) - m);
    const float g = scale * (cw ? cw[yy] : 1.f);
    const float inv = 1.f / sum;
    for (int k = 0; k < K; ++k) {
      const float sm = expf(__ldg(x + k * HW) - m) * inv;
      d[k * HW] = g * (sm - (k == yy ? 1.f : 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused top of the training ladder (led_head.py:101-146 -> decode_head.py:362-379 -> OhemCrossEntropy): the loss of
//   logits = resize(r1, (H, W), bilinear, align_corners=False)          r1: NHWC [N, h, w, K], the half-resolution rung
// WITHOUT the full-resolution logit tensor.  Round 2's launch list had the un-fused chain - resize forward (957 MB out at
// 12 x 1024^2 x 19), NHWC->NCHW, the OHEM kernels over NCHW, the gradient's NCHW->NHWC and the resize backward gather - at
// 17 ms of a 63 ms training step, all of it moving a tensor that only exists to be reduced.  Here every full-resolution
// pixel interpolates its K logits in registers (same fmaf sequence as resize_fwd, so the values are bit-identical), and
// the backward is a gather per r1 pixel over the outputs that read it, each re-deriving its softmax - the order of
// resize_bwd's sums is kept, so d(r1) is bit-identical to the un-fused path as well.
constexpr int kMaxK = 32;

template <int KT>
__device__ __forceinline__ void up_logits(const float* __restrict__ r1n, int y, int x, int h, int w, int K, float sh, float sw,
                                          float* v) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
  bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
  const float* a = r1n + ((int64_t)y0 * w + x0) * K;
  const float* b = r1n + ((int64_t)y0 * w + x1) * K;
  const float* c = r1n + ((int64_t)y1 * w + x0) * K;
  const float* d = r1n + ((int64_t)y1 * w + x1) * K;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k < K) {
      const float r0 = fmaf(__ldg(b + k), lx1, __ldg(a + k) * lx0);
      const float r1v = fmaf(__ldg(d + k), lx1, __ldg(c + k) * lx0);
      v[k] = fmaf(r1v, ly1, r0 * ly0);
    }
  }
}

template <int KT>
__global__ void __launch_bounds__(kT)
ohem_up_pixel_kernel(const float* __restrict__ r1, const int64_t* __restrict__ target, int K, int h, int w, int H, int W,
                     int64_t npix, int ignore, const float* __restrict__ cw, float sh, float sw, float* __restrict__ prob,
                     float* __restrict__ loss, OhemState* s) {
  __shared__ float shm[kT / 32];
  float nvalid = 0.f, ncorrect = 0.f;
  const int64_t HW = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < npix; i += (int64_t)gridDim.x * kT) {
    const int64_t n = i / HW, hw = i % HW;
    const int y = (int)(hw / W), x = (int)(hw % W);
    float v[KT];
    up_logits<KT>(r1 + n * (int64_t)h * w * K, y, x, h, w, K, sh, sw, v);
    const int64_t yt = target[i];
    float m = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K && v[k] > m) { m = v[k]; am = k; }
    float p = 2.0f, l = 0.f;
    if (yt != ignore) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) sum += expf(v[k] - m);
      const int yy = (yt >= 0 && yt < K) ? (int)yt : 0;
      float xy = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k == yy) xy = v[k] - m;
      p = expf(xy) / sum;
      l = -(xy - logf(sum)) * (cw ? cw[yy] : 1.f);
      nvalid += 1.f;
      ncorrect += (am == (int)yt) ? 1.f : 0.f;
    }
    prob[i] = p;
    loss[i] = l;
  }
  const float a = block_sum(nvalid, shm);
  const float b = block_sum(ncorrect, shm);
  if (threadIdx.x == 0) {
    atomicAdd(&s->nvalid, (unsigned long long)a);
    atomicAdd(&s->ncorrect, (unsigned long long)b);
  }
}

// candidate outputs that may read source index `sidx` (same bounds as train.cu's gather_range)
__device__ __forceinline__ void up_gather_range(int sidx, float scale, int out_size, int& lo, int& hi) {
  lo = (int)floorf(((float)sidx - 0.5f) / scale - 0.5f) - 1;
  hi = (int)ceilf(((float)sidx + 1.5f) / scale - 0.5f) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

// d(r1)[n, ys, xs, :] = sum over the outputs (y, x) whose bilinear footprint holds (ys, xs) of wy * wx * d(logits)[n, :, y, x],
// d(logits) = gscale * loss_weight / kept * (softmax - onehot) on kept pixels (ohem_backward_kernel), rows then columns ascending
template <int KT>
__global__ void __launch_bounds__(128)
ohem_up_bwd_kernel(const float* __restrict__ r1, const int64_t* __restrict__ target, const float* __restrict__ prob, int K,
                   int h, int w, int H, int W, int64_t nsrc, int ignore, const float* __restrict__ cw, float loss_weight,
                   const float* __restrict__ gscale, float sh, float sw, const OhemState* __restrict__ s,
                   float* __restrict__ dr1) {
  const float thr = s->threshold;
  const float scale = loss_weight / s->kept * (gscale ? gscale[0] : 1.f);
  const int64_t hwsrc = (int64_t)h * w;
  for (int64_t i = blockIdx.x * 128ll + threadIdx.x; i < nsrc; i += (int64_t)gridDim.x * 128) {
    const int64_t n = i / hwsrc, r = i % hwsrc;
    const int ys = (int)(r / w), xs = (int)(r % w);
    const float* r1n = r1 + n * hwsrc * K;
    int ylo, yhi, xlo, xhi;
    up_gather_range(ys, sh, H, ylo, yhi);
    up_gather_range(xs, sw, W, xlo, xhi);
    float acc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      int y0, y1;
      float ly0, ly1;
      bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
      if (y0 != ys && y1 != ys) continue;
      float wy = 0.f;
      if (y0 == ys) wy += ly0;
      if (y1 == ys) wy += ly1;
      float rowv[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) rowv[k] = 0.f;
      for (int x = xlo; x <= xhi; ++x) {
        int x0, x1;
        float lx0, lx1;
        bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
        if (x0 != xs && x1 != xs) continue;
        float wx = 0.f;
        if (x0 == xs) wx += lx0;
        if (x1 == xs) wx += lx1;
        const int64_t pi = (n * H + y) * W + x;
        const float p = prob[pi];
        if (!((p < thr) && (p <= 1.5f))) continue;          // not kept: its gradient row is zero (fmaf(0, wx, row) == row)
        float v[KT];
        up_logits<KT>(r1n, y, x, h, w, K, sh, sw, v);
        const int64_t yt = target[pi];
        const int yy = (yt >= 0 && yt < K) ? (int)yt : 0;
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < K) m = fmaxf(m, v[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < K) sum += expf(v[k] - m);
        const float g = scale * (cw ? cw[yy] : 1.f);
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < K) rowv[k] = fmaf(g * (expf(v[k] - m) * inv - (k == yy ? 1.f : 0.f)), wx, rowv[k]);
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) acc[k] = fmaf(rowv[k], wy, acc[k]);
    }
    float* o = dr1 + i * K;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) o[k] = acc[k];
  }
}

// Tiled forms.  Flat one-thread-per-pixel kernels read the four neighbour rungs with 4 x K scalar loads whose addresses are
// 76 bytes apart across a warp (K = 19): LSU-bound at ~10 x the HBM time.  A CTA instead stages the r1 patch its outputs read
// into shared memory with coalesced row copies (odd pixel stride: conflict-free), and interpolates from there.
//   forward : 32 x 32 outputs per CTA
//   backward: 8 x 16 r1 pixels per CTA; d(logits) of every output that reads one of them ((2*8+2) x (2*16+2) for the x2
//             ladder) is written to shared memory first - ONE softmax per output instead of one per (output, source) pair -
//             then each thread gathers its source pixel's footprint from there, in the same order as the flat kernel.
constexpr int kTileSY = 8, kTileSX = 16, kTileO = 32;

template <int KT>
__device__ __forceinline__ void up_logits_smem(const float* __restrict__ sp, int sc, int sy0, int sx0, int y, int x, int h, int w,
                                               int K, int KS, float sh, float sw, float* v) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
  bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
  const float* a = sp + ((y0 - sy0) * sc + (x0 - sx0)) * KS;
  const float* b = sp + ((y0 - sy0) * sc + (x1 - sx0)) * KS;
  const float* c = sp + ((y1 - sy0) * sc + (x0 - sx0)) * KS;
  const float* d = sp + ((y1 - sy0) * sc + (x1 - sx0)) * KS;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k < K) {
      const float r0 = fmaf(b[k], lx1, a[k] * lx0);
      const float r1v = fmaf(d[k], lx1, c[k] * lx0);
      v[k] = fmaf(r1v, ly1, r0 * ly0);
    }
  }
}

// stage r1[n, sy0..sy1, sx0..sx1, :] as [row][col][KS]; rows are contiguous in global memory
__device__ __forceinline__ void stage_patch(const float* __restrict__ r1n, float* __restrict__ sp, int w, int K, int KS, int sy0,
                                            int sy1, int sx0, int sx1, int nthr) {
  const int sc = sx1 - sx0 + 1, rowlen = sc * K;
  for (int r = sy0; r <= sy1; ++r) {
    const float* src = r1n + ((int64_t)r * w + sx0) * K;
    float* dst = sp + (r - sy0) * sc * KS;
    for (int i = threadIdx.x; i < rowlen; i += nthr) dst[(i / K) * KS + (i % K)] = __ldg(src + i);
  }
}

template <int KT, bool EXACT>
__global__ void __launch_bounds__(kT)
ohem_up_pixel_tiled_kernel(const float* __restrict__ r1, const int64_t* __restrict__ target, int K_, int h, int w, int H, int W,
                           int tiles_x, int tiles_y, int cap, int ignore, const float* __restrict__ cw, float sh, float sw,
                           float* __restrict__ prob, float* __restrict__ loss, OhemState* s) {
  extern __shared__ float sp[];
  __shared__ float shm[kT / 32];
  const int K = EXACT ? KT : K_;
  const int KS = K | 1;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, n = blockIdx.x / (tiles_x * tiles_y);
  const int oy0 = ty * kTileO, ox0 = tx * kTileO, oy1 = min(oy0 + kTileO, H) - 1, ox1 = min(ox0 + kTileO, W) - 1;
  int sy0, sy1, sx0, sx1, t0;
  float f0, f1;
  bilinear_coord(oy0, sh, h, sy0, t0, f0, f1);
  bilinear_coord(oy1, sh, h, t0, sy1, f0, f1);
  bilinear_coord(ox0, sw, w, sx0, t0, f0, f1);
  bilinear_coord(ox1, sw, w, t0, sx1, f0, f1);
  const int sc = sx1 - sx0 + 1;
  if ((sy1 - sy0 + 1) * sc * KS > cap) __trap();
  stage_patch(r1 + (int64_t)n * h * w * K, sp, w, K, KS, sy0, sy1, sx0, sx1, kT);
  __syncthreads();
  float nvalid = 0.f, ncorrect = 0.f;
  const int tw = ox1 - ox0 + 1, th = oy1 - oy0 + 1;
  for (int o = threadIdx.x; o < tw * th; o += kT) {
    const int y = oy0 + o / tw, x = ox0 + o % tw;
    const int64_t i = ((int64_t)n * H + y) * W + x;
    float v[KT];
    {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
      bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
      const float* a = sp + ((y0 - sy0) * sc + (x0 - sx0)) * KS;
      const float* b = sp + ((y0 - sy0) * sc + (x1 - sx0)) * KS;
      const float* c = sp + ((y1 - sy0) * sc + (x0 - sx0)) * KS;
      const float* d = sp + ((y1 - sy0) * sc + (x1 - sx0)) * KS;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (EXACT || k < K) {
          const float r0 = fmaf(b[k], lx1, a[k] * lx0);
          const float r1v = fmaf(d[k], lx1, c[k] * lx0);
          v[k] = fmaf(r1v, ly1, r0 * ly0);
        }
    }
    const int64_t yt = target[i];
    float m = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if ((EXACT || k < K) && v[k] > m) { m = v[k]; am = k; }
    float p = 2.0f, l = 0.f;
    if (yt != ignore) {
      const int yy = (yt >= 0 && yt < K) ? (int)yt : 0;
      float sum = 0.f, xy = 0.f, ey = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (EXACT || k < K) {
          const float e = expf(v[k] - m);          // expf costs ~10 instructions: evaluated once per class
          sum += e;
          if (k == yy) { xy = v[k] - m; ey = e; }
        }
      p = ey / sum;
      l = -(xy - logf(sum)) * (cw ? cw[yy] : 1.f);
      nvalid += 1.f;
      ncorrect += (am == (int)yt) ? 1.f : 0.f;
    }
    prob[i] = p;
    loss[i] = l;
  }
  const float a = block_sum(nvalid, shm);
  const float b = block_sum(ncorrect, shm);
  if (threadIdx.x == 0) {
    atomicAdd(&s->nvalid, (unsigned long long)a);
    atomicAdd(&s->ncorrect, (unsigned long long)b);
  }
}

// EXACT: K == KT (the class loops carry no `k < K` guards - a quarter of the instructions of the guarded form).
// 256 threads: all of them produce d(logits) of the region; in the gather two threads share a source pixel, one per half of
// the classes, so no warp idles through it.
template <int KT, bool EXACT>
__global__ void __launch_bounds__(2 * kTileSY * kTileSX)
ohem_up_bwd_tiled_kernel(const float* __restrict__ r1, const int64_t* __restrict__ target, const float* __restrict__ prob,
                         int K_, int h, int w, int H, int W, int tiles_x, int tiles_y, int cap_g, int cap_p, int ignore,
                         const float* __restrict__ cw, float loss_weight, const float* __restrict__ gscale, float sh, float sw,
                         const OhemState* __restrict__ s, float* __restrict__ dr1) {
  extern __shared__ float sg[];                 // [region outputs][KS] d(logits), then the staged r1 patch
  constexpr int NT = 2 * kTileSY * kTileSX;
  const int K = EXACT ? KT : K_;
  const int KS = K | 1;                         // odd pixel stride: conflict-free when lanes walk neighbouring pixels
  const float thr = s->threshold;
  const float scale = loss_weight / s->kept * (gscale ? gscale[0] : 1.f);
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, n = blockIdx.x / (tiles_x * tiles_y);
  const int ys0 = ty * kTileSY, xs0 = tx * kTileSX;
  const int ys1 = min(ys0 + kTileSY, h) - 1, xs1 = min(xs0 + kTileSX, w) - 1;
  int ylo, yhi, xlo, xhi, t0;
  up_gather_range(ys0, sh, H, ylo, t0);
  up_gather_range(ys1, sh, H, t0, yhi);
  up_gather_range(xs0, sw, W, xlo, t0);
  up_gather_range(xs1, sw, W, t0, xhi);
  const int rh = yhi - ylo + 1, rw = xhi - xlo + 1;
  // r1 patch read by the region's outputs
  int sy0, sy1, sx0, sx1;
  float f0, f1;
  bilinear_coord(ylo, sh, h, sy0, t0, f0, f1);
  bilinear_coord(yhi, sh, h, t0, sy1, f0, f1);
  bilinear_coord(xlo, sw, w, sx0, t0, f0, f1);
  bilinear_coord(xhi, sw, w, t0, sx1, f0, f1);
  const int sc = sx1 - sx0 + 1;
  if (rh * rw * KS > cap_g || (sy1 - sy0 + 1) * sc * KS > cap_p) __trap();
  float* sp = sg + cap_g;
  stage_patch(r1 + (int64_t)n * h * w * K, sp, w, K, KS, sy0, sy1, sx0, sx1, NT);
  // candidate outputs and weights of the tile's 8 source rows / 16 source columns, once per CTA (the gather below would
  // otherwise run a bilinear_coord per (row, column) candidate pair and thread: half of the kernel's instructions)
  __shared__ int t_lo[kTileSY + kTileSX], t_n[kTileSY + kTileSX];
  __shared__ float t_w[kTileSY + kTileSX][8];
  if (threadIdx.x < kTileSY + kTileSX) {
    const bool isrow = threadIdx.x < kTileSY;
    const int sidx = isrow ? ys0 + (int)threadIdx.x : xs0 + (int)threadIdx.x - kTileSY;
    const float sc_ = isrow ? sh : sw;
    const int in_sz = isrow ? h : w, out_sz = isrow ? H : W;
    int lo, hi;
    up_gather_range(sidx, sc_, out_sz, lo, hi);
    int cnt = hi - lo + 1;
    if (cnt > 8 || sidx >= in_sz) cnt = -1;                  // -1: walk the candidates on the fly
    t_lo[threadIdx.x] = lo; t_n[threadIdx.x] = cnt;
    for (int k = 0; k < 8; ++k) {
      float wv = -1.f;                                       // -1: not a member (a member's weight may be exactly 0)
      if (k < cnt) {
        int i0, i1;
        float l0, l1;
        bilinear_coord(lo + k, sc_, in_sz, i0, i1, l0, l1);
        if (i0 == sidx || i1 == sidx) {
          wv = 0.f;
          if (i0 == sidx) wv += l0;
          if (i1 == sidx) wv += l1;
        }
      }
      t_w[threadIdx.x][k] = wv;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < rh * rw; o += NT) {
    const int y = ylo + o / rw, x = xlo + o % rw;
    const int64_t pi = ((int64_t)n * H + y) * W + x;
    const float p = prob[pi];
    float* dst = sg + o * KS;
    if (!((p < thr) && (p <= 1.5f))) {
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (EXACT || k < K) dst[k] = 0.f;
      continue;
    }
    float v[KT];
    {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
      bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
      const float* a = sp + ((y0 - sy0) * sc + (x0 - sx0)) * KS;
      const float* b = sp + ((y0 - sy0) * sc + (x1 - sx0)) * KS;
      const float* c = sp + ((y1 - sy0) * sc + (x0 - sx0)) * KS;
      const float* d = sp + ((y1 - sy0) * sc + (x1 - sx0)) * KS;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (EXACT || k < K) {
          const float r0 = fmaf(b[k], lx1, a[k] * lx0);
          const float r1v = fmaf(d[k], lx1, c[k] * lx0);
          v[k] = fmaf(r1v, ly1, r0 * ly0);
        }
    }
    const int64_t yt = target[pi];
    const int yy = (yt >= 0 && yt < K) ? (int)yt : 0;
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (EXACT || k < K) m = fmaxf(m, v[k]);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (EXACT || k < K) { v[k] = expf(v[k] - m); sum += v[k]; }     // one expf per class (~10 instructions each)
    const float g = scale * (cw ? cw[yy] : 1.f);
    const float inv = 1.f / sum;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (EXACT || k < K) dst[k] = g * (v[k] * inv - (k == yy ? 1.f : 0.f));
  }
  __syncthreads();
  // gather: thread pair (2 p, 2 p + 1) owns source pixel p; the first takes classes [0, KH), the second [KH, K)
  constexpr int KH = (KT + 1) / 2;
  const int pidx = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int ys = ys0 + pidx / kTileSX, xs = xs0 + pidx % kTileSX;
  if (ys > ys1 || xs > xs1) return;
  const int k0 = half * KH;
  const int ti = pidx / kTileSX, tj = kTileSY + pidx % kTileSX;
  const bool tab = t_n[ti] >= 0 && t_n[tj] >= 0;
  int cylo, cyhi, cxlo, cxhi;
  if (tab) {
    cylo = t_lo[ti]; cyhi = cylo + t_n[ti] - 1; cxlo = t_lo[tj]; cxhi = cxlo + t_n[tj] - 1;
  } else {
    up_gather_range(ys, sh, H, cylo, cyhi);
    up_gather_range(xs, sw, W, cxlo, cxhi);
  }
  float acc[KH];
#pragma unroll
  for (int k = 0; k < KH; ++k) acc[k] = 0.f;
  for (int y = cylo; y <= cyhi; ++y) {
    float wy;
    if (tab) {
      wy = t_w[ti][y - cylo];
      if (wy < 0.f) continue;
    } else {
      int y0, y1;
      float ly0, ly1;
      bilinear_coord(y, sh, h, y0, y1, ly0, ly1);
      if (y0 != ys && y1 != ys) continue;
      wy = 0.f;
      if (y0 == ys) wy += ly0;
      if (y1 == ys) wy += ly1;
    }
    float rowv[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) rowv[k] = 0.f;
    for (int x = cxlo; x <= cxhi; ++x) {
      float wx;
      if (tab) {
        wx = t_w[tj][x - cxlo];
        if (wx < 0.f) continue;
      } else {
        int x0, x1;
        float lx0, lx1;
        bilinear_coord(x, sw, w, x0, x1, lx0, lx1);
        if (x0 != xs && x1 != xs) continue;
        wx = 0.f;
        if (x0 == xs) wx += lx0;
        if (x1 == xs) wx += lx1;
      }
      const float* src = sg + ((y - ylo) * rw + (x - xlo)) * KS + k0;
#pragma unroll
      for (int k = 0; k < KH; ++k)
        if (k0 + k < K) rowv[k] = fmaf(src[k], wx, rowv[k]);
    }
#pragma unroll
    for (int k = 0; k < KH; ++k) acc[k] = fmaf(rowv[k], wy, acc[k]);
  }
  float* o = dr1 + (((int64_t)n * h + ys) * w + xs) * K + k0;
#pragma unroll
  for (int k = 0; k < KH; ++k)
    if (k0 + k < K) o[k] = acc[k];
}

// host mirrors of the device index formulas, for sizing the shared-memory tiles (the kernels trap if a tile exceeds them)
void host_bilinear(int dst, float scale, int in_size, int& i0, int& i1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
}
void host_gather_range(int sidx, float scale, int out_size, int& lo, int& hi) {
  lo = (int)floorf(((float)sidx - 0.5f) / scale - 0.5f) - 1;
  hi = (int)ceilf(((float)sidx + 1.5f) / scale - 0.5f) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}
// backward: largest (output region, r1 patch) extents along one axis over all source tiles of `tile` pixels
void bwd_extents(int src, int out, float scale, int tile, int& region, int& patch) {
  region = patch = 0;
  for (int s0 = 0; s0 < src; s0 += tile) {
    const int s1 = std::min(s0 + tile, src) - 1;
    int lo, hi, t, p0, p1;
    host_gather_range(s0, scale, out, lo, t);
    host_gather_range(s1, scale, out, t, hi);
    host_bilinear(lo, scale, src, p0, t);
    host_bilinear(hi, scale, src, t, p1);
    region = std::max(region, hi - lo + 1);
    patch = std::max(patch, p1 - p0 + 1);
  }
}
// forward: largest r1 patch extent along one axis over all output tiles
int fwd_extent(int src, int out, float scale) {
  int best = 0;
  for (int o0 = 0; o0 < out; o0 += kTileO) {
    const int o1 = std::min(o0 + kTileO, out) - 1;
    int p0, p1, t;
    host_bilinear(o0, scale, src, p0, t);
    host_bilinear(o1, scale, src, t, p1);
    best = std::max(best, p1 - p0 + 1);
  }
  return best;
}

}  // namespace

int launch_ohem_up(const float* r1, const int64_t* target, int N, int K, int h, int w, int H, int W, int ignore_label,
                   float thres, int64_t min_kept, float loss_weight, const float* class_weight, float* out3,
                   void* workspace, cudaStream_t st) {
  if (K < 1 || K > kMaxK || N < 0 || H < 1 || W < 1 || h < 1 || w < 1) return fail(LEDB200_EINVAL, "ohem_up: bad shape (K <= 32)");
  if (!workspace) return fail(LEDB200_EINVAL, "ohem_up: workspace is null");
  const int64_t npix = (int64_t)N * H * W;
  auto* s = reinterpret_cast<OhemState*>(workspace);
  float* prob = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((sizeof(OhemState) + 255) / 256) * 256);
  float* loss = prob + npix;
  if (min_kept < 1) min_kept = 1;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 8, ceil_div64(npix, kT)));
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  ohem_init_kernel<<<1, 256, 0, st>>>(s);
  bool tiled = false;
  if (npix > 0) {
    const int cap = (fwd_extent(h, H, sh) + 1) * (fwd_extent(w, W, sw) + 1) * (K | 1);
    const int tiles_x = ceil_div(W, kTileO), tiles_y = ceil_div(H, kTileO);
    const int64_t tiles = (int64_t)N * tiles_x * tiles_y;
    if ((size_t)cap * 4 <= 96 * 1024 && tiles < (1ll << 31)) {
      tiled = true;
      auto run = [&](auto kern) -> int {
        LEDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        kern<<<(unsigned)tiles, kT, (size_t)cap * 4, st>>>(r1, target, K, h, w, H, W, tiles_x, tiles_y, cap, ignore_label,
                                                         class_weight, sh, sw, prob, loss, s);
        LEDB_LAUNCH_OK("ohem_up_pixel_tiled_kernel");
        return LEDB200_OK;
      };
      const int rc = K == 19 ? run(ohem_up_pixel_tiled_kernel<19, true>)
                     : K == 2 ? run(ohem_up_pixel_tiled_kernel<2, true>)
                     : K <= 8 ? run(ohem_up_pixel_tiled_kernel<8, false>)
                     : K <= 20 ? run(ohem_up_pixel_tiled_kernel<20, false>) : run(ohem_up_pixel_tiled_kernel<32, false>);
      if (rc) return rc;
    }
  }
  if (npix > 0 && !tiled) {
    if (K <= 8) ohem_up_pixel_kernel<8><<<grid, kT, 0, st>>>(r1, target, K, h, w, H, W, npix, ignore_label, class_weight, sh, sw, prob, loss, s);
    else if (K <= 20) ohem_up_pixel_kernel<20><<<grid, kT, 0, st>>>(r1, target, K, h, w, H, W, npix, ignore_label, class_weight, sh, sw, prob, loss, s);
    else ohem_up_pixel_kernel<32><<<grid, kT, 0, st>>>(r1, target, K, h, w, H, W, npix, ignore_label, class_weight, sh, sw, prob, loss, s);
  }
  ohem_rank_kernel<<<1, 1, 0, st>>>(s, (long long)min_kept);
  const int shifts[3] = {21, 10, 0}, nbins[3] = {2048, 2048, 1024};
  for (int p = 0; p < 3; ++p) {
    if (npix > 0) ohem_hist_kernel<<<grid, kT, 0, st>>>(prob, npix, s, p, shifts[p], nbins[p]);
    ohem_find_kernel<<<1, 1, 0, st>>>(s, p, shifts[p], nbins[p], thres);
  }
  ohem_reduce_kernel<<<kPartials, kT, 0, st>>>(prob, loss, npix, s);
  ohem_final_kernel<<<1, 1, 0, st>>>(s, loss_weight, out3);
  LEDB_LAUNCH_OK("ohem_up kernels");
  return LEDB200_OK;
}

int launch_ohem_up_bwd(const float* r1, const int64_t* target, int N, int K, int h, int w, int H, int W, int ignore_label,
                       float loss_weight, const float* class_weight, const float* gscale, const void* workspace, float* dr1,
                       cudaStream_t st) {
  if (K < 1 || K > kMaxK) return fail(LEDB200_EINVAL, "ohem_up_bwd: bad shape (K <= 32)");
  if (!workspace || !dr1) return fail(LEDB200_EINVAL, "ohem_up_bwd: null buffer");
  const int64_t nsrc = (int64_t)N * h * w;
  if (nsrc == 0) return LEDB200_OK;
  auto* s = reinterpret_cast<const OhemState*>(workspace);
  const float* prob = reinterpret_cast<const float*>(reinterpret_cast<const char*>(workspace) + ((sizeof(OhemState) + 255) / 256) * 256);
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  {
    int ry, py, rx, px;
    bwd_extents(h, H, sh, kTileSY, ry, py);
    bwd_extents(w, W, sw, kTileSX, rx, px);
    const int cap_g = (ry + 1) * (rx + 1) * (K | 1), cap_p = (py + 1) * (px + 1) * (K | 1);
    const size_t smem = (size_t)(cap_g + cap_p) * sizeof(float);
    const int tiles_x = ceil_div(w, kTileSX), tiles_y = ceil_div(h, kTileSY);
    const int64_t tiles = (int64_t)N * tiles_x * tiles_y;
    if (smem <= 160 * 1024 && tiles < (1ll << 31)) {
      auto run = [&](auto kern) -> int {
        LEDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        kern<<<(unsigned)tiles, 2 * kTileSY * kTileSX, smem, st>>>(r1, target, prob, K, h, w, H, W, tiles_x, tiles_y, cap_g, cap_p, ignore_label,
                                                  class_weight, loss_weight, gscale, sh, sw, s, dr1);
        LEDB_LAUNCH_OK("ohem_up_bwd_tiled_kernel");
        return LEDB200_OK;
      };
      if (K == 19) return run(ohem_up_bwd_tiled_kernel<19, true>);     // Cityscapes
      if (K == 2) return run(ohem_up_bwd_tiled_kernel<2, true>);       // the reference config's two-class set
      if (K <= 8) return run(ohem_up_bwd_tiled_kernel<8, false>);
      if (K <= 20) return run(ohem_up_bwd_tiled_kernel<20, false>);
      return run(ohem_up_bwd_tiled_kernel<32, false>);
    }
  }
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 16, ceil_div64(nsrc, 128)));
  if (K <= 8) ohem_up_bwd_kernel<8><<<grid, 128, 0, st>>>(r1, target, prob, K, h, w, H, W, nsrc, ignore_label, class_weight, loss_weight, gscale, sh, sw, s, dr1);
  else if (K <= 20) ohem_up_bwd_kernel<20><<<grid, 128, 0, st>>>(r1, target, prob, K, h, w, H, W, nsrc, ignore_label, class_weight, loss_weight, gscale, sh, sw, s, dr1);
  else ohem_up_bwd_kernel<32><<<grid, 128, 0, st>>>(r1, target, prob, K, h, w, H, W, nsrc, ignore_label, class_weight, loss_weight, gscale, sh, sw, s, dr1);
  LEDB_LAUNCH_OK("ohem_up_bwd_kernel");
  return LEDB200_OK;
}

namespace {
}  // namespace

int64_t ohem_workspace_bytes(int64_t npix) {
  return (int64_t)sizeof(OhemState) + 256 + 2 * npix * (int64_t)sizeof(float);
}

int launch_ohem(const float* logits, const int64_t* target, int N, int K, int H, int W, int ignore_label,
                float thres, int64_t min_kept, float loss_weight, const float* class_weight, float* out3,
                float* dlogits, void* workspace, cudaStream_t st) {
  if (K < 1 || N < 0 || H < 0 || W < 0) return fail(LEDB200_EINVAL, "ohem: bad shape");
  if (!workspace) return fail(LEDB200_EINVAL, "ohem: workspace is null");
  const int64_t HW = (int64_t)H * W, npix = HW * N;
  auto* s = reinterpret_cast<OhemState*>(workspace);
  float* prob = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((sizeof(OhemState) + 255) / 256) * 256);
  float* loss = prob + npix;
  if (min_kept < 1) min_kept = 1;   // ctor: max(1, min_kept)  (ohem_cross_entropy_loss.py:47)
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 8, ceil_div64(npix, kT)));
  ohem_init_kernel<<<1, 256, 0, st>>>(s);
  if (npix > 0) {
    ohem_pixel_kernel<<<grid, kT, 0, st>>>(logits, target, K, HW, npix, ignore_label, class_weight, prob, loss, s);
  }
  ohem_rank_kernel<<<1, 1, 0, st>>>(s, (long long)min_kept);
  const int shifts[3] = {21, 10, 0}, nbins[3] = {2048, 2048, 1024};
  for (int p = 0; p < 3; ++p) {
    if (npix > 0) ohem_hist_kernel<<<grid, kT, 0, st>>>(prob, npix, s, p, shifts[p], nbins[p]);
    ohem_find_kernel<<<1, 1, 0, st>>>(s, p, shifts[p], nbins[p], thres);
  }
  ohem_reduce_kernel<<<kPartials, kT, 0, st>>>(prob, loss, npix, s);
  ohem_final_kernel<<<1, 1, 0, st>>>(s, loss_weight, out3);
  if (dlogits && npix > 0)
    ohem_backward_kernel<<<grid, kT, 0, st>>>(logits, target, prob, K, HW, npix, ignore_label, class_weight,
                                              loss_weight, s, dlogits);
  LEDB_LAUNCH_OK("ohem kernels");
  return LEDB200_OK;
}

}  // namespace ledb
