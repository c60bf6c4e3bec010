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
