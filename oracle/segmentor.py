"""Oracle segmentor plumbing (TEST INFRASTRUCTURE).

Follows ``mmseg/models/segmentors/encoder_decoder.py:117-132,187-345`` (whole /
slide inference), ``mmseg/models/segmentors/base.py:127-200`` (postprocess_result:
un-pad, resize to ori_shape, argmax(dim=0)), ``mmseg/models/data_preprocessor.py:98-151``
and ``mmseg/utils/misc.py:30-128`` (stack_batch).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .r0 import OracleLEDNet, resize
from .head import OracleLEDHead
from .metrics import intersect_and_union

MEAN = (123.675, 116.28, 103.53)     # configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:14-15
STD = (58.395, 57.12, 57.375)


def preprocess(imgs_u8, mean=MEAN, std=STD, bgr_to_rgb=True):
    """data_preprocessor.py:112-118 on a uint8 [N,3,H,W] tensor (BGR in)."""
    x = imgs_u8
    if bgr_to_rgb:
        x = x[:, [2, 1, 0]]
    x = x.float()
    m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return (x - m) / s


def stack_batch(inputs, labels=None, size=None, size_divisor=None, pad_val=0,
                seg_pad_val=255):
    """misc.py:30-128: right/bottom pad every CHW tensor to ``size`` (or to the
    batch max rounded up to ``size_divisor``); labels padded with seg_pad_val."""
    assert (size is not None) ^ (size_divisor is not None)
    hs = max(t.shape[-2] for t in inputs)
    ws = max(t.shape[-1] for t in inputs)
    if size_divisor is not None and size_divisor > 1:
        hs = (hs + size_divisor - 1) // size_divisor * size_divisor
        ws = (ws + size_divisor - 1) // size_divisor * size_divisor
    out, out_l, pads = [], [], []
    for i, t in enumerate(inputs):
        th, tw = (size[-2], size[-1]) if size is not None else (hs, ws)
        pad = (0, max(tw - t.shape[-1], 0), 0, max(th - t.shape[-2], 0))
        out.append(F.pad(t, pad, value=pad_val))
        pads.append(pad)
        if labels is not None:
            out_l.append(F.pad(labels[i], pad, value=seg_pad_val))
    lab = torch.stack(out_l, 0) if labels is not None else None
    return torch.stack(out, 0), lab, pads


def postprocess_argmax(seg_logits, padding=None, ori_shape=None, align_corners=False):
    """base.py:153-198 for C > 1: crop padding, resize to ori_shape, argmax(dim=0).
    Returns int64 [N,1,h,w] (list-stacked; all images share a shape here)."""
    n, c, h, w = seg_logits.shape
    preds = []
    for i in range(n):
        lg = seg_logits[i:i + 1]
        if padding is not None:
            l, r, t, b = padding[i]
            lg = lg[:, :, t:h - b, l:w - r]
        if ori_shape is not None:
            lg = resize(lg, ori_shape, align_corners)
        preds.append(lg.squeeze(0).argmax(dim=0, keepdim=True))
    return torch.stack(preds, 0)


def slide_inference(encode_decode, inputs, num_classes, crop_size, stride):
    """encoder_decoder.py:241-292."""
    h_stride, w_stride = stride
    h_crop, w_crop = crop_size
    n, _, h_img, w_img = inputs.shape
    h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
    preds = inputs.new_zeros((n, num_classes, h_img, w_img))
    count = inputs.new_zeros((n, 1, h_img, w_img))
    for hi in range(h_grids):
        for wi in range(w_grids):
            y1, x1 = hi * h_stride, wi * w_stride
            y2, x2 = min(y1 + h_crop, h_img), min(x1 + w_crop, w_img)
            y1, x1 = max(y2 - h_crop, 0), max(x2 - w_crop, 0)
            logit = encode_decode(inputs[:, :, y1:y2, x1:x2])
            preds += F.pad(logit, (x1, w_img - x2, y1, h_img - y2))
            count[:, :, y1:y2, x1:x2] += 1
    assert (count == 0).sum() == 0
    return preds / count


class OracleSegmentor(nn.Module):
    """EncoderDecoder(LEDNet, LEDHead) restated: extract_feat -> decode_head.predict
    -> postprocess argmax -> IoUMetric.intersect_and_union."""

    def __init__(self, num_classes=2, channels=32, ppm_channels=128, head_channels=64,
                 align_corners=False, test_cfg=None):
        super().__init__()
        self.backbone = OracleLEDNet(3, channels, ppm_channels, align_corners)
        self.decode_head = OracleLEDHead(channels * 4, head_channels, num_classes,
                                         align_corners=align_corners,
                                         tap_channels=channels)
        self.num_classes = num_classes
        self.align_corners = align_corners
        self.test_cfg = test_cfg or dict(mode='whole')

    def encode_decode(self, x):
        return self.decode_head.predict(self.backbone(x))

    def inference(self, x):
        if self.test_cfg.get('mode', 'whole') == 'slide':
            return slide_inference(self.encode_decode, x, self.num_classes,
                                   self.test_cfg['crop_size'], self.test_cfg['stride'])
        return self.encode_decode(x)

    @torch.no_grad()
    def predict(self, x):
        """[N,3,H,W] float -> (seg_logits [N,K,H',W'], pred int64 [N,1,H',W'])."""
        logits = self.inference(x)
        return logits, postprocess_argmax(logits)

    @torch.no_grad()
    def predict_and_score(self, x, labels, ignore_index=255):
        logits, pred = self.predict(x)
        res = [intersect_and_union(pred[i, 0], labels[i], self.num_classes, ignore_index)
               for i in range(x.shape[0])]
        return logits, pred, res

    def loss(self, x, labels):
        return self.decode_head.loss(self.backbone(x), labels)
