// Result post-processing on the device.
//
//  * postprocess_kernel : BaseSegmentor.postprocess_result (mmseg/models/segmentors/base.py:153-198) for one
//      image: remove the padding border, undo a horizontal / vertical test-time flip, bilinear-resize the
//      logits to `ori_shape` (mmseg/models/utils/wrappers.py:8-27 -> F.interpolate, align_corners as the
//      decode head's) and take argmax(dim=0) (C > 1) or sigmoid > threshold (C == 1) - in ONE pass: the
//      resized [K, oh, ow] logits are written only when the caller asks for them.  The real Apple-Branch
//      pipeline needs the general ratio (512x910 -> 720x1280, scale 1.40625); all BASELINE configs are the
//      identity case, which the fused tail kernel already covers.
//  * slide_accumulate / slide_finalize : EncoderDecoder.slide_inference (encoder_decoder.py:241-292):
//      `preds += F.pad(crop_logits, ...)`, `count_mat[..., y1:y2, x1:x2] += 1` without materialising the padded
//      full-size tensor per crop, then `preds / count_mat` (+ optional argmax) in one pass.
//  * slide_merge : the same for ALL crops at once (they ran as one batch of the engine): per output pixel the crops that
//      cover it are summed in the reference's grid order (fp32, starting from 0: bit-identical to the sequential
//      `preds +=`), divided by their number, and optionally arg-maxed - no full-size read-modify-write per crop.
//  * stack_pad : SegDataPreProcessor.forward + stack_batch for one sample (mmseg/models/data_preprocessor.py:112-149,
//      mmseg/utils/misc.py:30-128): BGR<->RGB swap, float, (x - mean) / std, right / bottom padding with `pad_val`
//      (applied AFTER the normalisation, as the reference does), label map padded with `seg_pad_val`.
// All tensors are the reference's NCHW fp32; one thread per output pixel, consecutive threads on consecutive x.
#include <vector>

#include "kernels.h"

namespace ledb {
namespace {

struct PostArgs {
  const float* logits;   // [K, H, W] of one image
  int K, H, W;
  int pl, pr, pt, pb;    // padding to remove
  int flip;              // 0 none, 1 horizontal, 2 vertical
  int oh, ow, align_corners;
  float sy, sx;          // ATen scales
  float threshold;       // C == 1
  void* pred; int pred_dtype;
  float* out_logits;     // optional [K, oh, ow]
};

__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int align, int& i0, int& i1, float& l1) {
  if (align) {
    const float src = scale * (float)dst;                         // area_pixel_compute_source_index, align_corners
    i0 = min((int)src, in_size - 1);
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - (float)i0;
  } else {
    float l0;
    bilinear_coord(dst, scale, in_size, i0, i1, l0, l1);
  }
}

template <typename TP>
__global__ void __launch_bounds__(256) postprocess_kernel(PostArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)a.oh * a.ow) return;
  const int ox = (int)(idx % a.ow), oy = (int)(idx / a.ow);
  const int ch = a.H - a.pt - a.pb, cw = a.W - a.pl - a.pr;       // cropped extent
  int y0, y1, x0, x1;
  float ly, lx;
  src_index(oy, a.sy, ch, a.align_corners, y0, y1, ly);
  src_index(ox, a.sx, cw, a.align_corners, x0, x1, lx);
  // cropped (and flipped) coordinates -> coordinates of the stored logits
  if (a.flip == 1) { x0 = cw - 1 - x0; x1 = cw - 1 - x1; }
  if (a.flip == 2) { y0 = ch - 1 - y0; y1 = ch - 1 - y1; }
  const int64_t o00 = (int64_t)(y0 + a.pt) * a.W + x0 + a.pl, o01 = (int64_t)(y0 + a.pt) * a.W + x1 + a.pl;
  const int64_t o10 = (int64_t)(y1 + a.pt) * a.W + x0 + a.pl, o11 = (int64_t)(y1 + a.pt) * a.W + x1 + a.pl;
  const float hy = 1.f - ly, hx = 1.f - lx;
  float best = -INFINITY;
  int bi = 0;
  float v = 0.f;
  for (int k = 0; k < a.K; ++k) {
    const float* p = a.logits + (int64_t)k * a.H * a.W;
    // ATen upsample_bilinear2d: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
    v = hy * (hx * __ldg(p + o00) + lx * __ldg(p + o01)) + ly * (hx * __ldg(p + o10) + lx * __ldg(p + o11));
    if (a.K == 1) v = 1.f / (1.f + __expf(-v));
    if (a.out_logits) a.out_logits[(int64_t)k * a.oh * a.ow + idx] = v;
    if (v > best) { best = v; bi = k; }                           // strict >: torch.argmax's first-max rule
  }
  TP* pred = reinterpret_cast<TP*>(a.pred);
  if (pred) pred[idx] = (TP)(a.K == 1 ? (v > a.threshold ? 1 : 0) : bi);
}

__global__ void __launch_bounds__(256) slide_accumulate_kernel(float* preds, float* count, const float* crop, int N, int K,
                                                               int H, int W, int hc, int wc, int y1, int x1) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)N * (K + 1) * hc * wc;            // plane K of every image = the count matrix
  if (idx >= total) return;
  const int x = (int)(idx % wc);
  int64_t t = idx / wc;
  const int y = (int)(t % hc); t /= hc;
  const int k = (int)(t % (K + 1));
  const int n = (int)(t / (K + 1));
  const int64_t pix = (int64_t)(y1 + y) * W + x1 + x;
  if (k < K) preds[((int64_t)n * K + k) * H * W + pix] += crop[(((int64_t)n * K + k) * hc + y) * wc + x];
  else count[(int64_t)n * H * W + pix] += 1.f;
}

template <typename TP>
__global__ void __launch_bounds__(256) slide_finalize_kernel(float* preds, const float* count, int N, int K, int64_t HW,
                                                             TP* pred) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * HW) return;
  const int n = (int)(idx / HW);
  const int64_t pix = idx % HW;
  const float c = count[idx];
  float best = -INFINITY;
  int bi = 0;
  for (int k = 0; k < K; ++k) {
    float* p = preds + ((int64_t)n * K + k) * HW + pix;
    const float v = *p / c;                                        // preds / count_mat (encoder_decoder.py:290)
    *p = v;
    if (v > best) { best = v; bi = k; }
  }
  if (pred) pred[idx] = (TP)bi;
}

constexpr int kMaxCrops = 64;
struct MergeArgs {
  const float* crops;    // [G * N, K, hc, wc], crop g of image n at index g * N + n
  int G, N, K, H, W, hc, wc;
  int y1[kMaxCrops], x1[kMaxCrops];
  float* out;            // optional [N, K, H, W]
  void* pred;            // optional [N, H, W]
};

template <typename TP>
__global__ void __launch_bounds__(256) slide_merge_kernel(const __grid_constant__ MergeArgs a) {
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)a.N * HW) return;
  const int n = (int)(idx / HW);
  const int64_t pix = idx % HW;
  const int y = (int)(pix / a.W), x = (int)(pix % a.W);
  const int64_t plane = (int64_t)a.hc * a.wc;
  float cnt = 0.f;
  for (int g = 0; g < a.G; ++g)
    cnt += (y >= a.y1[g] && y < a.y1[g] + a.hc && x >= a.x1[g] && x < a.x1[g] + a.wc) ? 1.f : 0.f;
  float best = -INFINITY;
  int bi = 0;
  for (int k = 0; k < a.K; ++k) {
    float s = 0.f;                                               // preds = zeros; preds += pad(crop) in grid order
    for (int g = 0; g < a.G; ++g) {
      const int yy = y - a.y1[g], xx = x - a.x1[g];
      if (yy >= 0 && yy < a.hc && xx >= 0 && xx < a.wc)
        s += __ldg(a.crops + (((int64_t)g * a.N + n) * a.K + k) * plane + (int64_t)yy * a.wc + xx);
    }
    const float v = s / cnt;
    if (a.out) a.out[((int64_t)n * a.K + k) * HW + pix] = v;
    if (v > best) { best = v; bi = k; }
  }
  if (a.pred) reinterpret_cast<TP*>(a.pred)[idx] = (TP)bi;
}

struct StackPadArgs {
  const void* img; int h, w, swap_rb, normalize;
  float mean[3], stdv[3], pad_val;
  float* out; int Hp, Wp;
  const void* label; int64_t* label_out; int seg_pad_val;
};

template <typename TI, typename TL>
__global__ void __launch_bounds__(256) stack_pad_kernel(const __grid_constant__ StackPadArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t HWp = (int64_t)a.Hp * a.Wp;
  if (idx >= HWp) return;
  const int y = (int)(idx / a.Wp), x = (int)(idx % a.Wp);
  const bool inside = y < a.h && x < a.w;
  const int64_t src = (int64_t)y * a.w + x, hw = (int64_t)a.h * a.w;
  const TI* img = reinterpret_cast<const TI*>(a.img);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = a.pad_val;
    if (inside) {
      v = (float)img[(a.swap_rb ? 2 - c : c) * hw + src];
      if (a.normalize) v = __fdiv_rn(v - a.mean[c], a.stdv[c]);   // (_input - self.mean) / self.std, channel c AFTER the swap
    }
    a.out[c * HWp + idx] = v;
  }
  if (a.label_out)
    a.label_out[idx] = inside ? (int64_t)reinterpret_cast<const TL*>(a.label)[src] : (int64_t)a.seg_pad_val;
}

}  // namespace
}  // namespace ledb

using namespace ledb;

extern "C" {

int ledb200_postprocess(const float* logits, int32_t K, int32_t H, int32_t W, const int32_t* padding_lrtb,
                        int32_t flip, int32_t out_h, int32_t out_w, int32_t align_corners, float threshold,
                        void* pred, int32_t pred_dtype, float* out_logits, void* stream) {
  if (!logits || (!pred && !out_logits)) return fail(LEDB200_EINVAL, "postprocess: null buffer");
  if (K < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return fail(LEDB200_EINVAL, "postprocess: empty tensor");
  if (flip < 0 || flip > 2) return fail(LEDB200_EINVAL, "postprocess: flip must be 0 (none), 1 (horizontal) or 2 (vertical)");
  if (pred && pred_dtype != LEDB200_U8 && pred_dtype != LEDB200_I64 && pred_dtype != LEDB200_F32)
    return fail(LEDB200_EINVAL, "postprocess: pred dtype must be U8, I64 or F32");
  PostArgs a;
  a.logits = logits; a.K = K; a.H = H; a.W = W;
  a.pl = padding_lrtb ? padding_lrtb[0] : 0; a.pr = padding_lrtb ? padding_lrtb[1] : 0;
  a.pt = padding_lrtb ? padding_lrtb[2] : 0; a.pb = padding_lrtb ? padding_lrtb[3] : 0;
  const int ch = H - a.pt - a.pb, cw = W - a.pl - a.pr;
  if (a.pl < 0 || a.pr < 0 || a.pt < 0 || a.pb < 0 || ch < 1 || cw < 1)
    return fail(LEDB200_EINVAL, "postprocess: padding removes the whole image");
  a.flip = flip; a.oh = out_h; a.ow = out_w; a.align_corners = align_corners ? 1 : 0;
  if (a.align_corners) {
    a.sy = out_h > 1 ? (float)(ch - 1) / (float)(out_h - 1) : 0.f;
    a.sx = out_w > 1 ? (float)(cw - 1) / (float)(out_w - 1) : 0.f;
  } else {
    a.sy = (float)ch / (float)out_h; a.sx = (float)cw / (float)out_w;
  }
  a.threshold = threshold; a.pred = pred; a.pred_dtype = pred_dtype; a.out_logits = out_logits;
  const unsigned grid = (unsigned)ceil_div64((int64_t)out_h * out_w, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (pred_dtype == LEDB200_I64) postprocess_kernel<int64_t><<<grid, 256, 0, st>>>(a);
  else if (pred_dtype == LEDB200_F32) postprocess_kernel<float><<<grid, 256, 0, st>>>(a);
  else postprocess_kernel<uint8_t><<<grid, 256, 0, st>>>(a);
  LEDB_LAUNCH_OK("postprocess_kernel");
  return LEDB200_OK;
}

int ledb200_slide_accumulate(float* preds, float* count, const float* crop_logits, int32_t N, int32_t K, int32_t H,
                             int32_t W, int32_t hc, int32_t wc, int32_t y1, int32_t x1, void* stream) {
  if (!preds || !count || !crop_logits) return fail(LEDB200_EINVAL, "slide_accumulate: null buffer");
  if (N < 1 || K < 1 || hc < 1 || wc < 1 || y1 < 0 || x1 < 0 || y1 + hc > H || x1 + wc > W)
    return fail(LEDB200_EINVAL, "slide_accumulate: crop window outside the image");
  const int64_t total = (int64_t)N * (K + 1) * hc * wc;
  slide_accumulate_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(preds, count, crop_logits, N, K,
                                                                                          H, W, hc, wc, y1, x1);
  LEDB_LAUNCH_OK("slide_accumulate_kernel");
  return LEDB200_OK;
}

int ledb200_slide_finalize(float* preds, const float* count, int32_t N, int32_t K, int32_t H, int32_t W, void* pred,
                           int32_t pred_dtype, void* stream) {
  if (!preds || !count) return fail(LEDB200_EINVAL, "slide_finalize: null buffer");
  if (pred && pred_dtype != LEDB200_U8 && pred_dtype != LEDB200_I64)
    return fail(LEDB200_EINVAL, "slide_finalize: pred dtype must be U8 or I64");
  const int64_t HW = (int64_t)H * W;
  const unsigned grid = (unsigned)ceil_div64((int64_t)N * HW, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (pred_dtype == LEDB200_I64) slide_finalize_kernel<int64_t><<<grid, 256, 0, st>>>(preds, count, N, K, HW, (int64_t*)pred);
  else slide_finalize_kernel<uint8_t><<<grid, 256, 0, st>>>(preds, count, N, K, HW, (uint8_t*)pred);
  LEDB_LAUNCH_OK("slide_finalize_kernel");
  return LEDB200_OK;
}

int ledb200_slide_merge(const float* crop_logits, int32_t G, const int32_t* y1, const int32_t* x1, int32_t N, int32_t K,
                        int32_t H, int32_t W, int32_t hc, int32_t wc, float* out_logits, void* pred, int32_t pred_dtype,
                        void* stream) {
  if (!crop_logits || !y1 || !x1 || (!out_logits && !pred)) return fail(LEDB200_EINVAL, "slide_merge: null buffer");
  if (G < 1 || G > kMaxCrops) return fail(LEDB200_EINVAL, "slide_merge: between 1 and 64 crops per call");
  if (N < 1 || K < 1 || hc < 1 || wc < 1 || hc > H || wc > W) return fail(LEDB200_EINVAL, "slide_merge: bad shape");
  if (pred && pred_dtype != LEDB200_U8 && pred_dtype != LEDB200_I64)
    return fail(LEDB200_EINVAL, "slide_merge: pred dtype must be U8 or I64");
  MergeArgs a;
  a.crops = crop_logits; a.G = G; a.N = N; a.K = K; a.H = H; a.W = W; a.hc = hc; a.wc = wc; a.out = out_logits; a.pred = pred;
  // encoder_decoder.py:289 asserts count_mat != 0 everywhere.  The reference's windows are a cartesian product of row
  // and column offsets, so the image is covered iff every row and every column lies in some window: checked here.
  std::vector<char> rows(H, 0), cols(W, 0);
  for (int g = 0; g < G; ++g) {
    if (y1[g] < 0 || x1[g] < 0 || y1[g] + hc > H || x1[g] + wc > W)
      return fail(LEDB200_EINVAL, "slide_merge: crop window outside the image");
    a.y1[g] = y1[g]; a.x1[g] = x1[g];
  }
  for (int g = 0; g < G; ++g) {
    for (int y = y1[g]; y < y1[g] + hc; ++y) rows[y] = 1;
    for (int x = x1[g]; x < x1[g] + wc; ++x) cols[x] = 1;
  }
  for (int y = 0; y < H; ++y) if (!rows[y]) return fail(LEDB200_EINVAL, "slide_merge: the crop windows leave rows uncovered");
  for (int x = 0; x < W; ++x) if (!cols[x]) return fail(LEDB200_EINVAL, "slide_merge: the crop windows leave columns uncovered");
  const unsigned grid = (unsigned)ceil_div64((int64_t)N * H * W, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (pred_dtype == LEDB200_I64) slide_merge_kernel<int64_t><<<grid, 256, 0, st>>>(a);
  else slide_merge_kernel<uint8_t><<<grid, 256, 0, st>>>(a);
  LEDB_LAUNCH_OK("slide_merge_kernel");
  return LEDB200_OK;
}

int ledb200_stack_pad(const void* img, int32_t img_dtype, int32_t h, int32_t w, int32_t swap_rb, const float* mean3,
                      const float* std3, float pad_val, float* out, int32_t Hp, int32_t Wp, const void* label,
                      int32_t label_dtype, int64_t* label_out, int32_t seg_pad_val, void* stream) {
  if (!img || !out) return fail(LEDB200_EINVAL, "stack_pad: null buffer");
  if (h < 1 || w < 1 || Hp < h || Wp < w) return fail(LEDB200_EINVAL, "stack_pad: the padded size is smaller than the image");
  if (img_dtype != LEDB200_U8 && img_dtype != LEDB200_F32) return fail(LEDB200_EINVAL, "stack_pad: image dtype must be U8 or F32");
  if ((mean3 == nullptr) != (std3 == nullptr)) return fail(LEDB200_EINVAL, "stack_pad: mean and std come together");
  if ((label == nullptr) != (label_out == nullptr)) return fail(LEDB200_EINVAL, "stack_pad: label and label_out come together");
  if (label && label_dtype != LEDB200_U8 && label_dtype != LEDB200_I64)
    return fail(LEDB200_EINVAL, "stack_pad: label dtype must be U8 or I64");
  StackPadArgs a;
  a.img = img; a.h = h; a.w = w; a.swap_rb = swap_rb ? 1 : 0; a.normalize = mean3 ? 1 : 0;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean3 ? mean3[c] : 0.f; a.stdv[c] = std3 ? std3[c] : 1.f; }
  a.pad_val = pad_val; a.out = out; a.Hp = Hp; a.Wp = Wp; a.label = label; a.label_out = label_out; a.seg_pad_val = seg_pad_val;
  const unsigned grid = (unsigned)ceil_div64((int64_t)Hp * Wp, 256);
  cudaStream_t st = (cudaStream_t)stream;
  const bool l64 = label && label_dtype == LEDB200_I64;
  if (img_dtype == LEDB200_U8) {
    if (l64) stack_pad_kernel<uint8_t, int64_t><<<grid, 256, 0, st>>>(a);
    else stack_pad_kernel<uint8_t, uint8_t><<<grid, 256, 0, st>>>(a);
  } else {
    if (l64) stack_pad_kernel<float, int64_t><<<grid, 256, 0, st>>>(a);
    else stack_pad_kernel<float, uint8_t><<<grid, 256, 0, st>>>(a);
  }
  LEDB_LAUNCH_OK("stack_pad_kernel");
  return LEDB200_OK;
}

}  // extern "C"
