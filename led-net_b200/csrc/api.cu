// Stand-alone C-ABI operators (no handle): the fused head tail, the confusion matrix, OHEM CE and a
// single convolution through the same launchers the engine uses.  See include/ledb200.h.
#include <vector>

#include "kernels.h"

using namespace ledb;

extern "C" {

int ledb200_head_fuse_argmax(const void* xc, const void* hx2, const void* hx1, int32_t dtype, int32_t N, int32_t K,
                             int32_t hc, int32_t wc, int32_t h4, int32_t w4, int32_t h2, int32_t w2, void* pred,
                             int32_t pred_dtype, float* logits_opt, void* stream) {
  if (!xc || !hx2 || !hx1 || !pred) return fail(LEDB200_EINVAL, "head_fuse_argmax: null buffer");
  TailArgs a;
  a.xc = xc; a.hx2 = hx2; a.hx1 = hx1; a.xc_ld = a.hx2_ld = a.hx1_ld = K; a.dtype = dtype;
  a.N = N; a.K = K; a.hc = hc; a.wc = wc; a.h4 = h4; a.w4 = w4; a.h2 = h2; a.w2 = w2;
  a.pred = pred; a.pred_dtype = pred_dtype; a.logits = logits_opt;
  return launch_tail(a, (cudaStream_t)stream);
}

int ledb200_confusion_accumulate(const void* pred, const void* gt, int32_t pred_dtype, int32_t gt_dtype, int64_t n,
                                 int32_t K, int32_t ignore_index, int64_t* cm_inout, void* stream) {
  if ((!pred || !gt) && n > 0) return fail(LEDB200_EINVAL, "confusion: null buffer");
  if (!cm_inout) return fail(LEDB200_EINVAL, "confusion: null matrix");
  return launch_confusion(pred, gt, pred_dtype, gt_dtype, n, K, ignore_index, cm_inout, (cudaStream_t)stream);
}

int64_t ledb200_ohem_workspace_bytes(int64_t npix) { return ohem_workspace_bytes(npix); }

int ledb200_ohem_up_fwd(const float* r1_nhwc, const int64_t* target, int32_t N, int32_t K, int32_t h, int32_t w, int32_t H,
                        int32_t W, int32_t ignore_label, float thres, int64_t min_kept, float loss_weight,
                        const float* class_weight_opt, float* out3, void* workspace, void* stream) {
  if (!out3) return fail(LEDB200_EINVAL, "ohem_up: null output");
  if ((!r1_nhwc || !target) && (int64_t)N * H * W > 0) return fail(LEDB200_EINVAL, "ohem_up: null input");
  return launch_ohem_up(r1_nhwc, target, N, K, h, w, H, W, ignore_label, thres, min_kept, loss_weight, class_weight_opt, out3,
                        workspace, (cudaStream_t)stream);
}

int ledb200_ohem_up_bwd(const float* r1_nhwc, const int64_t* target, int32_t N, int32_t K, int32_t h, int32_t w, int32_t H,
                        int32_t W, int32_t ignore_label, float loss_weight, const float* class_weight_opt,
                        const float* grad_scale_opt, const void* workspace, float* d_r1, void* stream) {
  if (!r1_nhwc || !target) return fail(LEDB200_EINVAL, "ohem_up_bwd: null input");
  return launch_ohem_up_bwd(r1_nhwc, target, N, K, h, w, H, W, ignore_label, loss_weight, class_weight_opt, grad_scale_opt,
                            workspace, d_r1, (cudaStream_t)stream);
}

int ledb200_ohem_ce(const float* logits, const int64_t* target, int32_t N, int32_t K, int32_t H, int32_t W,
                    int32_t ignore_label, float thres, int64_t min_kept, float loss_weight,
                    const float* class_weight_opt, float* out3, float* dlogits_opt, void* workspace, void* stream) {
  if (!out3) return fail(LEDB200_EINVAL, "ohem: null output");
  if ((!logits || !target) && (int64_t)N * H * W > 0) return fail(LEDB200_EINVAL, "ohem: null input");
  return launch_ohem(logits, target, N, K, H, W, ignore_label, thres, min_kept, loss_weight, class_weight_opt, out3,
                     dlogits_opt, workspace, (cudaStream_t)stream);
}

int ledb200_conv2d(const void* in, void* out, const void* residual, int32_t dtype, int32_t N, int32_t H, int32_t W,
                   int32_t Cin, int32_t Cout, int32_t ksize, int32_t stride, int32_t relu, const float* weight_oihw,
                   const float* bias, const float* pre_scale, const float* pre_shift, int32_t backend, int32_t in_ld,
                   int32_t out_ld, int32_t res_ld, void* stream) {
  if (!in || !out || !weight_oihw) return fail(LEDB200_EINVAL, "conv2d: null buffer");
  if (ksize != 1 && ksize != 3) return fail(LEDB200_EINVAL, "conv2d: ksize must be 1 or 3");
  if (stride != 1 && stride != 2) return fail(LEDB200_EINVAL, "conv2d: stride must be 1 or 2");
  if (dtype != LEDB200_F32 && dtype != LEDB200_BF16) return fail(LEDB200_EINVAL, "conv2d: dtype must be F32 or BF16");
  if (in_ld == 0) in_ld = Cin;
  if (out_ld == 0) out_ld = Cout;
  if (res_ld == 0) res_ld = Cout;
  if (in_ld < Cin || out_ld < Cout || res_ld < Cout) return fail(LEDB200_EINVAL, "conv2d: pixel stride smaller than the channel count");
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = ksize * ksize, cp16 = (Cout + 15) / 16 * 16, cptc = conv_tc_pad(Cout);
  std::vector<float> wd((size_t)taps * Cin * cp16, 0.f), bz(cptc > cp16 ? cptc : cp16, 0.f);
  std::vector<__nv_bfloat16> wt((size_t)cptc * taps * Cin, __float2bfloat16(0.f));
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) {
        const float v = weight_oihw[((size_t)o * Cin + c) * taps + t];
        wd[((size_t)t * Cin + c) * cp16 + o] = v;
        wt[((size_t)o * taps + t) * Cin + c] = __float2bfloat16(v);
      }
  if (bias) for (int o = 0; o < Cout; ++o) bz[o] = bias[o];
  float *d_wd = nullptr, *d_b = nullptr, *d_ps = nullptr, *d_pb = nullptr;
  __nv_bfloat16* d_wt = nullptr;
  int rc = LEDB200_OK;
  auto up = [&](void** dst, const void* src, size_t bytes) -> int {
    LEDB_CUDA_OK(cudaMalloc(dst, bytes));
    LEDB_CUDA_OK(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st));
    return LEDB200_OK;
  };
  rc = up((void**)&d_wd, wd.data(), wd.size() * 4);
  if (!rc) rc = up((void**)&d_wt, wt.data(), wt.size() * 2);
  if (!rc) rc = up((void**)&d_b, bz.data(), bz.size() * 4);
  if (!rc && pre_scale) rc = up((void**)&d_ps, pre_scale, Cin * 4);
  if (!rc && pre_shift) rc = up((void**)&d_pb, pre_shift, Cin * 4);
  if (!rc) {
    ConvArgs a;
    const int pad = ksize / 2;
    a.in = in; a.in_dtype = dtype; a.in_sc = 1; a.in_sw = in_ld; a.in_sh = (int64_t)W * in_ld; a.in_sn = (int64_t)H * W * in_ld;
    a.out = out; a.out_dtype = dtype; a.out_ld = out_ld; a.res = residual; a.res_ld = res_ld;
    a.bias = d_b; a.pre_scale = d_ps; a.pre_shift = d_pb; a.pre_relu = 1;
    a.w_direct = d_wd; a.w_tc = d_wt; a.cout_pad16 = cp16; a.cout_pad_tc = cptc;
    a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.stride = stride; a.pad = pad; a.dil = 1;
    a.Ho = (H + 2 * pad - ksize) / stride + 1; a.Wo = (W + 2 * pad - ksize) / stride + 1; a.relu = relu;
    if (backend == 2) {
      if (dtype != LEDB200_BF16 || pre_scale || !conv_tc_eligible(a))
        rc = fail(LEDB200_EINVAL, "conv2d: shape not eligible for the tcgen05 path");
      else rc = launch_conv_tc(a, st);
    } else if (backend == 0 && dtype == LEDB200_BF16 && !pre_scale && conv_tc_eligible(a)) {
      rc = launch_conv_tc(a, st);
    } else {
      rc = launch_conv_direct(a, st);
    }
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  if (!rc && ce != cudaSuccess) rc = fail(LEDB200_ECUDA, std::string("conv2d: ") + cudaGetErrorString(ce));
  cudaFree(d_wd); cudaFree(d_wt); cudaFree(d_b); cudaFree(d_ps); cudaFree(d_pb);
  return rc;
}

}  // extern "C"
