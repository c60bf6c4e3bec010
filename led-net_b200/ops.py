"""Functional wrappers over the stand-alone C-ABI kernels (include/ledb200.h).

Layout changes between the reference's NCHW tensors and the library's NHWC are views/copies made
with torch (plumbing); the arithmetic is in the CUDA kernels.
"""
import ctypes as C

import torch

from . import lib as L


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.LedB200Error('LED-Net B200 ops need CUDA tensors (no CPU fallback)')


def head_fuse_argmax(xc, hx2, hx1, pred_dtype=torch.uint8, want_logits=False, channels_last=False):
    """predict_by_feat + argmax (decode_head.py:362-379, base.py:187-188).

    xc/hx2/hx1: [N,K,h,w] (NCHW, default) or [N,h,w,K] (channels_last=True), fp32 or bf16.
    Returns (pred [N,2*h2,2*w2], logits or None)."""
    _need_cuda(xc, hx2, hx1)
    if not channels_last:
        xc, hx2, hx1 = (t.permute(0, 2, 3, 1).contiguous() for t in (xc, hx2, hx1))
    else:
        xc, hx2, hx1 = (t.contiguous() for t in (xc, hx2, hx1))
    assert xc.dtype == hx2.dtype == hx1.dtype and xc.dtype in (torch.float32, torch.bfloat16)
    N, hc, wc, K = xc.shape
    _, h4, w4, _ = hx2.shape
    _, h2, w2, _ = hx1.shape
    pred = torch.empty((N, 2 * h2, 2 * w2), dtype=pred_dtype, device=xc.device)
    logits = torch.empty((N, K, 2 * h2, 2 * w2), dtype=torch.float32, device=xc.device) if want_logits else None
    L.check(L.get().ledb200_head_fuse_argmax(
        _p(xc), _p(hx2), _p(hx1), L.torch_dtype_code(xc), N, K, hc, wc, h4, w4, h2, w2, _p(pred),
        L.torch_dtype_code(pred), _p(logits), L.stream_ptr(xc.device)), 'ledb200_head_fuse_argmax')
    return pred, logits


def confusion_accumulate(pred, gt, num_classes, ignore_index=255, cm=None):
    """cm[(K+1),K] int64 += bincount(K*gt+pred) over gt != ignore (rows GT; row K = out-of-range GT)."""
    _need_cuda(pred, gt)
    pred, gt = pred.contiguous(), gt.contiguous()
    assert pred.numel() == gt.numel(), 'prediction and label must have the same number of pixels'
    if cm is None:
        cm = torch.zeros((num_classes + 1, num_classes), dtype=torch.int64, device=pred.device)
    L.check(L.get().ledb200_confusion_accumulate(
        _p(pred), _p(gt), L.torch_dtype_code(pred), L.torch_dtype_code(gt), pred.numel(), num_classes,
        ignore_index, _p(cm), L.stream_ptr(pred.device)), 'ledb200_confusion_accumulate')
    return cm


def ohem_ce(score, target, ignore_label=255, thres=0.7, min_kept=100000, loss_weight=1.0,
            class_weight=None, want_grad=False):
    """Returns (out3 = [loss, kept, accuracy%] device tensor, dlogits or None)."""
    _need_cuda(score, target)
    score = score.contiguous().float()
    target = target.contiguous().to(torch.int64)
    N, K, H, W = score.shape
    lib = L.get()
    ws = torch.empty(lib.ledb200_ohem_workspace_bytes(N * H * W), dtype=torch.uint8, device=score.device)
    out3 = torch.empty(3, dtype=torch.float32, device=score.device)
    grad = torch.empty_like(score) if want_grad else None
    cw = None
    if class_weight is not None:
        cw = torch.as_tensor(class_weight, dtype=torch.float32, device=score.device).contiguous()
        assert cw.numel() == K
    L.check(lib.ledb200_ohem_ce(_p(score), _p(target), N, K, H, W, ignore_label, float(thres),
                                int(min_kept), float(loss_weight), _p(cw), _p(out3), _p(grad), _p(ws),
                                L.stream_ptr(score.device)), 'ledb200_ohem_ce')
    return out3, grad


def conv2d(x_nhwc, weight_oihw, bias=None, stride=1, relu=False, residual=None, pre_scale=None,
           pre_shift=None, backend=0):
    """One convolution (3x3 pad 1 or 1x1) on an NHWC fp32/bf16 CUDA tensor; weights host fp32."""
    _need_cuda(x_nhwc, residual)
    x = x_nhwc.contiguous()
    N, H, W, Cin = x.shape
    w = weight_oihw.detach().to('cpu', torch.float32).contiguous()
    Cout, _, k, _ = w.shape
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    # 16-byte pixels for the tensor-core path: Cout 19 -> pixel stride 24, like the engine's buffers
    ld = (Cout + 7) // 8 * 8
    out_full = torch.empty((N, Ho, Wo, ld), dtype=x.dtype, device=x.device)
    host = [None if t is None else t.detach().to('cpu', torch.float32).contiguous()
            for t in (bias, pre_scale, pre_shift)]
    res = None
    if residual is not None:
        res = residual.new_zeros((N, Ho, Wo, ld))
        res[..., :Cout] = residual
    L.check(L.get().ledb200_conv2d(_p(x), _p(out_full), _p(res), L.torch_dtype_code(x), N, H, W, Cin, Cout, k,
                                   stride, int(relu), _p(w), _p(host[0]), _p(host[1]), _p(host[2]),
                                   backend, Cin, ld, ld, L.stream_ptr(x.device)), 'ledb200_conv2d')
    return out_full[..., :Cout]


def postprocess(seg_logit, padding=None, flip=None, ori_shape=None, align_corners=False, threshold=0.3,
                pred_dtype=torch.int64, want_logits=True):
    """BaseSegmentor.postprocess_result for ONE image (base.py:153-198): seg_logit fp32 [K,H,W] -> crop `padding`
    (left, right, top, bottom), undo `flip` ('horizontal' | 'vertical' | None), bilinear resize to `ori_shape`,
    argmax (K > 1) or sigmoid > threshold (K == 1).  Returns (pred [1,h,w], logits [K,h,w] or None)."""
    _need_cuda(seg_logit)
    x = seg_logit.contiguous().float()
    K, H, W = x.shape
    pad = tuple(int(v) for v in (padding if padding is not None else (0, 0, 0, 0)))
    assert flip in (None, False, 'horizontal', 'vertical'), flip
    oh, ow = (int(v) for v in ori_shape) if ori_shape is not None else (H - pad[2] - pad[3], W - pad[0] - pad[1])
    if K == 1:
        pred_dtype = torch.float32                    # (sigmoid > threshold).to(seg_logits) in the reference
    pred = torch.empty((1, oh, ow), dtype=pred_dtype, device=x.device)
    out = torch.empty((K, oh, ow), dtype=torch.float32, device=x.device) if want_logits else None
    pad_c = (C.c_int32 * 4)(*pad)
    L.check(L.get().ledb200_postprocess(_p(x), K, H, W, pad_c, {None: 0, False: 0, 'horizontal': 1, 'vertical': 2}[flip],
                                        oh, ow, int(bool(align_corners)), float(threshold), _p(pred),
                                        L.torch_dtype_code(pred), _p(out), L.stream_ptr(x.device)), 'ledb200_postprocess')
    return pred, out


def slide_accumulate(preds, count, crop_logits, y1, x1):
    """preds[:, :, y1:y1+hc, x1:x1+wc] += crop_logits; count[:, :, same] += 1 (encoder_decoder.py:283-287)."""
    _need_cuda(preds, count, crop_logits)
    assert preds.is_contiguous() and count.is_contiguous() and preds.dtype == count.dtype == torch.float32
    crop = crop_logits.contiguous().float()
    N, K, H, W = preds.shape
    _, _, hc, wc = crop.shape
    L.check(L.get().ledb200_slide_accumulate(_p(preds), _p(count), _p(crop), N, K, H, W, hc, wc, int(y1), int(x1),
                                             L.stream_ptr(preds.device)), 'ledb200_slide_accumulate')


def slide_finalize(preds, count, want_pred=False, pred_dtype=torch.int64):
    """preds /= count in place (encoder_decoder.py:290); optionally also argmax over K -> [N,H,W]."""
    _need_cuda(preds, count)
    N, K, H, W = preds.shape
    pred = torch.empty((N, H, W), dtype=pred_dtype, device=preds.device) if want_pred else None
    L.check(L.get().ledb200_slide_finalize(_p(preds), _p(count), N, K, H, W, _p(pred),
                                           L.torch_dtype_code(pred) if want_pred else L.I64,
                                           L.stream_ptr(preds.device)), 'ledb200_slide_finalize')
    return preds, pred


def slide_merge(crop_logits, origins, n_images, image_hw, want_logits=True, want_pred=False, pred_dtype=torch.int64):
    """All crop windows of slide_inference at once: crop_logits fp32 [G*N,K,hc,wc] (window g of image n at g*N+n),
    origins = [(y1, x1)] * G in the reference's grid order -> (logits [N,K,H,W] or None, pred [N,H,W] or None)."""
    _need_cuda(crop_logits)
    crop = crop_logits.contiguous().float()
    G, N = len(origins), int(n_images)
    GN, K, hc, wc = crop.shape
    assert GN == G * N, 'crop batch must hold every window of every image'
    H, W = (int(v) for v in image_hw)
    out = torch.empty((N, K, H, W), dtype=torch.float32, device=crop.device) if want_logits else None
    pred = torch.empty((N, H, W), dtype=pred_dtype, device=crop.device) if want_pred else None
    y1 = (C.c_int32 * G)(*[int(o[0]) for o in origins])
    x1 = (C.c_int32 * G)(*[int(o[1]) for o in origins])
    L.check(L.get().ledb200_slide_merge(_p(crop), G, y1, x1, N, K, H, W, hc, wc, _p(out), _p(pred),
                                        L.torch_dtype_code(pred) if want_pred else L.I64,
                                        L.stream_ptr(crop.device)), 'ledb200_slide_merge')
    return out, pred


def stack_pad(img, out, swap_rb=False, mean=None, std=None, pad_val=0.0, label=None, label_out=None, seg_pad_val=255):
    """SegDataPreProcessor + stack_batch for one sample: img [3,h,w] uint8 / fp32 -> out [3,Hp,Wp] fp32 (a slice of the
    batch tensor), label [h,w] or [1,h,w] uint8 / int64 -> label_out [Hp,Wp] int64."""
    _need_cuda(img, out, label, label_out)
    img = img.contiguous()
    assert img.dim() == 3 and img.shape[0] == 3 and out.is_contiguous() and out.dtype == torch.float32
    if img.dtype not in (torch.uint8, torch.float32):
        img = img.float()
    if label is not None:
        label = label.reshape(label.shape[-2:]).contiguous()
        if label.dtype not in (torch.uint8, torch.int64):
            label = label.to(torch.int64)
        assert label_out.is_contiguous() and label_out.dtype == torch.int64
    m3 = (C.c_float * 3)(*[float(v) for v in mean]) if mean is not None else None
    s3 = (C.c_float * 3)(*[float(v) for v in std]) if std is not None else None
    L.check(L.get().ledb200_stack_pad(_p(img), L.torch_dtype_code(img), img.shape[1], img.shape[2], int(bool(swap_rb)),
                                      m3, s3, float(pad_val), _p(out), out.shape[-2], out.shape[-1], _p(label),
                                      L.torch_dtype_code(label) if label is not None else L.U8, _p(label_out),
                                      int(seg_pad_val), L.stream_ptr(img.device)), 'ledb200_stack_pad')
    return out
