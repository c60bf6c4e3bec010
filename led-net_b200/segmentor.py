"""`EncoderDecoder` and `SegDataPreProcessor` (MODELS) - the callers either side of the hot path.

Reference surface: ``mmseg/models/segmentors/encoder_decoder.py:117-132, 187-345``,
``mmseg/models/segmentors/base.py:127-200`` and ``mmseg/models/data_preprocessor.py:98-151``.
`predict_labels` is the fused fast path (one C-ABI call: image batch -> label map, full-resolution
logits never materialised); `predict` keeps the reference's return shape (seg_logits + pred_sem_seg
per image) and therefore does materialise the logits.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import Engine, MEAN, STD
from .registry import MODELS
from . import ops


@MODELS.register_module()
class SegDataPreProcessor(nn.Module):
    """Holds mean/std/bgr_to_rgb.  On the fused path these are folded into the stem convolution's
    prologue (raw uint8 in); `forward` is the stand-alone torch version for callers that want the
    normalised float batch (plumbing, data_preprocessor.py:112-149, test-time branch)."""

    def __init__(self, mean=MEAN, std=STD, size=None, size_divisor=None, pad_val=0, seg_pad_val=255,
                 bgr_to_rgb=False, rgb_to_bgr=False, batch_augments=None, test_cfg=None):
        super().__init__()
        assert not (bgr_to_rgb and rgb_to_bgr)
        self.channel_conversion = bgr_to_rgb or rgb_to_bgr
        self.size, self.size_divisor, self.pad_val, self.seg_pad_val = size, size_divisor, pad_val, seg_pad_val
        self.test_cfg = test_cfg
        self.register_buffer('mean', torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1), False)
        self.register_buffer('std', torch.tensor(std, dtype=torch.float32).view(-1, 1, 1), False)

    def forward(self, data, training=False):
        inputs = data['inputs']
        if isinstance(inputs, (list, tuple)):
            inputs = torch.stack(list(inputs), 0)
        if self.channel_conversion and inputs.shape[1] == 3:
            inputs = inputs[:, [2, 1, 0]]
        inputs = (inputs.float() - self.mean) / self.std
        if training and self.size is not None:
            ph, pw = max(self.size[0] - inputs.shape[-2], 0), max(self.size[1] - inputs.shape[-1], 0)
            inputs = F.pad(inputs, (0, pw, 0, ph), value=self.pad_val)
        return dict(inputs=inputs, data_samples=data.get('data_samples'))


@MODELS.register_module()
class EncoderDecoder(nn.Module):

    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None,
                 test_cfg=None, data_preprocessor=None, pretrained=None, init_cfg=None,
                 compute_dtype='bf16'):
        super().__init__()
        assert neck is None and auxiliary_head is None, 'the LED-Net config has neither'
        self.backbone = MODELS.build(backbone) if isinstance(backbone, dict) else backbone
        self.decode_head = MODELS.build(decode_head) if isinstance(decode_head, dict) else decode_head
        self.data_preprocessor = (MODELS.build(data_preprocessor) if isinstance(data_preprocessor, dict)
                                  else data_preprocessor)
        self.align_corners = self.decode_head.align_corners       # encoder_decoder.py:103-105
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self.train_cfg, self.test_cfg = train_cfg, dict(test_cfg or dict(mode='whole'))
        self.compute_dtype = compute_dtype
        self._engine = None
        for part in (self.backbone, self.decode_head):          # stand-alone module engines (mode='tensor') follow
            if hasattr(part, 'set_compute_dtype'):
                part.set_compute_dtype(compute_dtype)
        self.register_load_state_dict_post_hook(lambda m, keys: m.reset_engine())

    # -- engine ---------------------------------------------------------------------------
    def reset_engine(self):
        self._engine = None

    def set_compute_dtype(self, dtype):
        if dtype != self.compute_dtype:
            self.compute_dtype, self._engine = dtype, None
        for part in (self.backbone, self.decode_head):
            if hasattr(part, 'set_compute_dtype'):
                part.set_compute_dtype(dtype)
        return self

    def engine(self):
        if self._engine is None:
            state = {'backbone.' + k: v for k, v in self.backbone.state_dict().items()}
            state.update({'decode_head.' + k: v for k, v in self.decode_head.state_dict().items()})
            pp = self.data_preprocessor
            kw = {}
            if pp is not None:
                kw = dict(mean=tuple(pp.mean.flatten().tolist()), std=tuple(pp.std.flatten().tolist()),
                          bgr_to_rgb=pp.channel_conversion)
            self._engine = Engine(state, self.num_classes, self.backbone.channels,
                                  self.backbone.ppm_channels, self.decode_head.channels,
                                  dtype=self.compute_dtype, **kw)
        return self._engine

    # -- reference API --------------------------------------------------------------------
    def extract_feat(self, inputs):
        if self.training:
            return self.backbone(inputs)            # train-mode tape over csrc/train.cu
        return self.engine().backbone_forward(inputs)

    def train(self, mode=True):
        if mode:
            self.reset_engine()                     # folded eval weights go stale once training resumes
        return super().train(mode)

    def loss(self, inputs, data_samples):
        """encoder_decoder.py:161-185: decode-head losses under the 'decode.' prefix."""
        x = self.extract_feat(inputs)
        out = self.decode_head.loss(x, data_samples, self.train_cfg)
        return {'decode.' + k: v for k, v in out.items()}

    @staticmethod
    def parse_losses(losses):
        """mmengine BaseModel.parse_losses: total = sum of every entry whose key contains 'loss'."""
        log = {k: (v.mean() if isinstance(v, torch.Tensor) else sum(x.mean() for x in v))
               for k, v in losses.items()}
        total = sum(v for k, v in log.items() if 'loss' in k)
        log['loss'] = total
        return total, log

    def train_step(self, data, optimizer):
        """mmengine BaseModel.train_step: preprocess -> loss -> backward -> optimiser step.
        `optimizer` is a lednet_b200.optim.FlatSGD (or any object with zero_grad()/step())."""
        if self.data_preprocessor is not None and isinstance(data, dict) and 'inputs' in data:
            data = self.data_preprocessor(data, training=True)
        total, log = self.parse_losses(self.loss(data['inputs'], data['data_samples']))
        optimizer.zero_grad()
        total.backward()
        optimizer.step()
        return log

    def encode_decode(self, inputs, batch_img_metas=None):
        """encoder_decoder.py:124-132: full-resolution logits [N,K,H,W] (fused, one call)."""
        return self.engine().forward_infer(inputs, want_logits=True)[1]

    def whole_inference(self, inputs, batch_img_metas=None):
        return self.encode_decode(inputs, batch_img_metas)

    def slide_inference(self, inputs, batch_img_metas=None):
        """encoder_decoder.py:241-292.  Every crop runs the fused engine; `preds += F.pad(...)`, the count matrix and
        the final division are the slide_accumulate / slide_finalize kernels (no padded full-size copy per crop)."""
        h_stride, w_stride = self.test_cfg['stride']
        h_crop, w_crop = self.test_cfg['crop_size']
        n, _, h_img, w_img = inputs.shape
        h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
        w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
        preds = torch.zeros((n, self.out_channels, h_img, w_img), dtype=torch.float32, device=inputs.device)
        count = torch.zeros((n, 1, h_img, w_img), dtype=torch.float32, device=inputs.device)
        for hi in range(h_grids):
            for wi in range(w_grids):
                y1, x1 = hi * h_stride, wi * w_stride
                y2, x2 = min(y1 + h_crop, h_img), min(x1 + w_crop, w_img)
                y1, x1 = max(y2 - h_crop, 0), max(x2 - w_crop, 0)
                logit = self.encode_decode(inputs[:, :, y1:y2, x1:x2].contiguous())
                ops.slide_accumulate(preds, count, logit, y1, x1)
        assert (count == 0).sum() == 0
        return ops.slide_finalize(preds, count)[0]

    def inference(self, inputs, batch_img_metas=None):
        mode = self.test_cfg.get('mode', 'whole')
        assert mode in ['slide', 'whole'], \
            f'Only "slide" or "whole" test mode are supported, but got {mode}.'
        return self.slide_inference(inputs, batch_img_metas) if mode == 'slide' \
            else self.whole_inference(inputs, batch_img_metas)

    def postprocess_result(self, seg_logits, data_samples=None):
        """base.py:153-198: per image remove the padding, undo the test-time flip, resize to `ori_shape`, argmax
        (or sigmoid > threshold for one class) - one kernel per image (ops.postprocess).  `data_samples` are dicts
        whose 'metainfo' (or the dict itself) carries img_padding_size / padding_size, flip, flip_direction, ori_shape."""
        out = []
        for i in range(seg_logits.shape[0]):
            meta = {}
            if data_samples is not None and isinstance(data_samples[i], dict):
                meta = data_samples[i].get('metainfo', data_samples[i])
            padding = meta.get('img_padding_size', meta.get('padding_size', [0] * 4))
            flip = None
            if meta.get('flip', None):
                flip = meta.get('flip_direction', None)
                assert flip in ['horizontal', 'vertical']
            pred, lg = ops.postprocess(seg_logits[i], padding=padding, flip=flip, ori_shape=meta.get('ori_shape'),
                                       align_corners=self.align_corners,
                                       threshold=getattr(self.decode_head, 'threshold', None) or 0.3)
            out.append({'seg_logits': {'data': lg}, 'pred_sem_seg': {'data': pred}})
        return out

    def predict(self, inputs, data_samples=None):
        """encoder_decoder.py:187-222: list of {'seg_logits','pred_sem_seg'} per image."""
        def _plain(s):                       # no crop / flip / resize requested for this sample
            m = s.get('metainfo', s) if isinstance(s, dict) else {}
            return not (m.get('flip') or m.get('ori_shape') is not None and tuple(m['ori_shape']) != tuple(inputs.shape[-2:])
                        or any(m.get('img_padding_size', m.get('padding_size', [0] * 4))))
        if self.test_cfg.get('mode', 'whole') == 'whole' and (data_samples is None or all(_plain(s) for s in data_samples)):
            pred, logits = self.engine().forward_infer(inputs, pred_dtype=torch.int64, want_logits=True)
            res = [{'seg_logits': {'data': logits[i]}, 'pred_sem_seg': {'data': pred[i:i + 1]}}
                   for i in range(pred.shape[0])]
        else:
            res = self.postprocess_result(self.inference(inputs), data_samples)
        if data_samples is not None:
            for r, s in zip(res, data_samples):
                if isinstance(s, dict):
                    r.update({k: v for k, v in s.items() if k not in r})
        return res

    @torch.no_grad()
    def predict_labels(self, inputs, pred=None, pred_dtype=torch.uint8):
        """Fused fast path: [N,3,H,W] float (normalised) or uint8 (raw BGR) -> labels [N,H,W]."""
        return self.engine().forward_infer(inputs, pred=pred, pred_dtype=pred_dtype)

    def forward(self, inputs, data_samples=None, mode='tensor'):
        if mode == 'predict':
            return self.predict(inputs, data_samples)
        if mode == 'tensor':
            return self.decode_head.forward(self.extract_feat(inputs))
        if mode == 'loss':
            return self.loss(inputs, data_samples)
        raise RuntimeError(f'Invalid mode "{mode}". Only supports loss, predict and tensor mode')
