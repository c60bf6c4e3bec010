"""The callers either side of the hot path on the device (SURVEY 8f rows f2, f3 + the binary head of VERDICT r1 item 7):
SegDataPreProcessor / stack_batch padding, batched slide_inference, `out_channels=1` heads.  Every call goes through the
C ABI (ops.stack_pad / ops.slide_merge / the engine)."""
import os
import warnings

import numpy as np
import pytest
import torch

import lednet_b200 as L
from lednet_b200 import ops, synth
import oracle
from block_cases import STACK_CASES, stack_inputs
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


@pytest.mark.parametrize('ci', range(len(STACK_CASES)))
def test_preprocessor_training_matches_reference_stack_batch(ci):
    """Golden: the reference's own stack_batch behind data_preprocessor.py:118-123 (tests/golden/make_golden.py stack).
    Bit-exact: (x - mean) / std is one IEEE subtraction and one division on both sides."""
    tag, shapes, size, div, pad_val, seg_pad_val = STACK_CASES[ci]
    gold = np.load(os.path.join(GOLD, 'stack.npz'))
    imgs, labs = stack_inputs(ci, shapes)
    pp = L.SegDataPreProcessor(mean=MEAN, std=STD, bgr_to_rgb=True, size=size, size_divisor=div, pad_val=pad_val,
                               seg_pad_val=seg_pad_val)
    samples = [dict(gt_sem_seg=dict(data=lab.to(DEV))) for lab in labs]
    out = pp(dict(inputs=[im.to(DEV) for im in imgs], data_samples=samples), training=True)
    np.testing.assert_array_equal(out['inputs'].cpu().numpy(), gold[tag + '_inputs'])
    got_lab = torch.stack([s['gt_sem_seg']['data'] for s in out['data_samples']]).cpu().numpy()
    np.testing.assert_array_equal(got_lab, gold[tag + '_labels'].astype(np.int64))
    for s, (h, w), pad in zip(out['data_samples'], shapes, gold[tag + '_padding']):
        assert tuple(s['metainfo']['padding_size']) == tuple(int(v) for v in pad)
        assert s['metainfo']['img_shape'] == (h, w)
        assert tuple(s['metainfo']['pad_shape']) == tuple(out['inputs'].shape[-2:])


def test_preprocessor_uint8_label_and_no_normalisation():
    """mean/std None = no normalisation (data_preprocessor.py:88-97); uint8 label maps are accepted."""
    imgs, labs = stack_inputs(7, [(19, 23), (19, 23)])
    pp = L.SegDataPreProcessor(rgb_to_bgr=True, size_divisor=8, pad_val=3, seg_pad_val=200)
    samples = [dict(gt_sem_seg=dict(data=lab.to(torch.uint8).to(DEV))) for lab in labs]
    out = pp(dict(inputs=torch.stack(imgs).to(DEV), data_samples=samples), training=True)
    ref = torch.nn.functional.pad(torch.stack(imgs)[:, [2, 1, 0]].float(), (0, 1, 0, 5), value=3.)
    assert torch.equal(out['inputs'].cpu(), ref)
    ref_lab = torch.nn.functional.pad(torch.stack(labs), (0, 1, 0, 5), value=200)
    assert torch.equal(torch.stack([s['gt_sem_seg']['data'] for s in samples]).cpu(), ref_lab)


def test_preprocessor_test_cfg_padding_round_trip():
    """Test-time padding (data_preprocessor.py:136-147): img_padding_size lands in the sample's metainfo and
    postprocess_result (base.py:163-176) removes exactly that border again."""
    img = synth.make_images_u8(2, 50, 70, seed=3)
    pp = L.SegDataPreProcessor(mean=MEAN, std=STD, bgr_to_rgb=True, test_cfg=dict(size_divisor=32))
    samples = [dict(), dict()]
    out = pp(dict(inputs=img.to(DEV), data_samples=samples), training=False)
    assert tuple(out['inputs'].shape) == (2, 3, 64, 96)
    assert samples[0]['img_padding_size'] == (0, 26, 0, 14) and tuple(samples[0]['pad_shape']) == (64, 96)
    ref = torch.nn.functional.pad(oracle.preprocess(img), (0, 26, 0, 14), value=0.)
    assert torch.equal(out['inputs'].cpu(), ref)
    # plain test-time stack (no test_cfg): normalisation only
    out2 = L.SegDataPreProcessor(mean=MEAN, std=STD, bgr_to_rgb=True)(dict(inputs=list(img.to(DEV))), training=False)
    assert torch.equal(out2['inputs'].cpu(), oracle.preprocess(img))


def test_preprocessor_errors():
    pp = L.SegDataPreProcessor(size=(32, 32))
    with pytest.raises(AssertionError):        # padded shapes differ: torch.stack would fail in the reference
        pp(dict(inputs=[torch.zeros(3, 40, 20, device=DEV), torch.zeros(3, 20, 20, device=DEV)],
                data_samples=[dict(), dict()]), training=True)
    with pytest.raises(AssertionError):        # only one of size / size_divisor (misc.py:64-66)
        L.SegDataPreProcessor(size=(32, 32), size_divisor=8)(
            dict(inputs=[torch.zeros(3, 8, 8, device=DEV)], data_samples=[dict()]), training=True)
    with pytest.raises(AssertionError):        # test-time batches hold one image size
        L.SegDataPreProcessor()(dict(inputs=[torch.zeros(3, 8, 8, device=DEV), torch.zeros(3, 8, 9, device=DEV)]))
    with pytest.raises(L.LedB200Error):
        ops.stack_pad(torch.zeros(3, 8, 8, device=DEV), torch.zeros(3, 4, 8, device=DEV))


@pytest.mark.parametrize('n,K,hw,crop,stride', [
    (2, 5, (70, 90), (32, 48), (20, 30)),      # ragged last windows are shifted back (encoder_decoder.py:272-275)
    (1, 19, (64, 128), (64, 64), (48, 48)),    # Cityscapes-shaped: one row of three windows
    (3, 2, (40, 40), (40, 40), (40, 40)),      # a single window
])
def test_slide_merge_equals_sequential_accumulate(n, K, hw, crop, stride):
    """ONE merge kernel over all windows == the reference's sequential `preds += pad(crop)`; `/= count`; argmax,
    bit for bit (same fp32 accumulation order per pixel)."""
    H, W = hw
    hc, wc = crop
    wins = []
    for hi in range(max(H - hc + stride[0] - 1, 0) // stride[0] + 1):
        for wi in range(max(W - wc + stride[1] - 1, 0) // stride[1] + 1):
            y2, x2 = min(hi * stride[0] + hc, H), min(wi * stride[1] + wc, W)
            wins.append((max(y2 - hc, 0), max(x2 - wc, 0)))
    g = torch.Generator().manual_seed(5)
    crops = torch.randn(len(wins) * n, K, hc, wc, generator=g)
    preds, count = torch.zeros(n, K, H, W), torch.zeros(n, 1, H, W)
    for gi, (y1, x1) in enumerate(wins):                        # encoder_decoder.py:283-290 on the CPU
        preds += torch.nn.functional.pad(crops[gi * n:(gi + 1) * n], (x1, W - x1 - wc, y1, H - y1 - hc))
        count[:, :, y1:y1 + hc, x1:x1 + wc] += 1
    ref = preds / count
    out, pred = ops.slide_merge(crops.to(DEV), wins, n, (H, W), want_logits=True, want_pred=True)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(pred.cpu(), ref.argmax(dim=1))
    # the per-window kernels agree too
    p2 = torch.zeros(n, K, H, W, device=DEV)
    c2 = torch.zeros(n, 1, H, W, device=DEV)
    for gi, (y1, x1) in enumerate(wins):
        ops.slide_accumulate(p2, c2, crops[gi * n:(gi + 1) * n].to(DEV), y1, x1)
    assert torch.equal(ops.slide_finalize(p2, c2)[0].cpu(), ref)


def test_slide_merge_rejects_uncovered_image():
    crops = torch.zeros(2, 3, 8, 8, device=DEV)
    with pytest.raises(L.LedB200Error):
        ops.slide_merge(crops, [(0, 0), (0, 8)], 1, (8, 20))    # columns 16..19 lie in no window


def _pair(K, **head_kw):
    o = oracle.OracleSegmentor(num_classes=K).eval()
    sd = synth.make_state_dict(o.state_dict(), seed=2)
    o.load_state_dict(sd)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64, num_classes=K,
                                                       dropout_ratio=0., **head_kw),
                             data_preprocessor=None, compute_dtype='fp32').eval()
    return o, m, sd


def test_slide_inference_batched_and_sequential_match_oracle():
    o, m, sd = _pair(2)
    m.load_state_dict(sd)
    o.test_cfg = dict(mode='slide', crop_size=(64, 64), stride=(48, 48))
    m.test_cfg = dict(o.test_cfg)
    x = oracle.preprocess(synth.make_images_u8(2, 96, 128, seed=11))
    ref = o.inference(x)
    got = m.inference(x.to(DEV))                                 # all 6 windows x 2 images as one engine batch
    assert rel_err(got.cpu(), ref) < 1e-4
    m.MAX_SLIDE_BATCH = 0                                        # per-window path
    got_seq = m.inference(x.to(DEV))
    assert rel_err(got_seq.cpu(), ref) < 1e-4
    assert rel_err(got_seq.cpu(), got.cpu()) < 1e-5              # batch size does not change a window's logits beyond fp32 noise
    _, pred = m.slide_inference(x.to(DEV), want_pred=True)
    assert torch.equal(pred.cpu(), got.cpu().argmax(dim=1))


def test_binary_head_out_channels_1_broadcasts_like_the_reference():
    """decode_head.py:119-138 + predict_by_feat (:362-379): with out_channels=1 the one-channel classifier map is
    broadcast-added to the 2-channel tap heads, so the result has 2 channels and postprocess takes the argmax branch.
    Check against the oracle head run with exactly that broadcast."""
    K = 2
    o, m, sd = _pair(K, out_channels=1)
    assert m.decode_head.threshold == 0.3 and m.decode_head.conv_seg.weight.shape[0] == 1
    sd1 = dict(sd)
    for name in ('decode_head.conv_seg', 'decode_head.aux_cls_seg'):
        sd1[name + '.weight'] = sd[name + '.weight'][:1].clone()
        sd1[name + '.bias'] = sd[name + '.bias'][:1].clone()
    m.load_state_dict(sd1)
    x = oracle.preprocess(synth.make_images_u8(2, 64, 96, seed=4))
    with torch.no_grad():
        c5, x1, x2 = o.backbone(x)
        h = o.decode_head
        xc1 = torch.nn.functional.conv2d(h.head(c5), sd1['decode_head.conv_seg.weight'], sd1['decode_head.conv_seg.bias'])
        ref = oracle.fuse_logits(xc1, h.head_x1(x1), h.head_x2(x2))      # [N,1,..] broadcast against [N,2,..]
    res = m.predict(x.to(DEV))
    got = torch.stack([r['seg_logits']['data'] for r in res]).cpu()
    assert got.shape[1] == K and rel_err(got, ref) < 1e-4
    pred = torch.stack([r['pred_sem_seg']['data'] for r in res]).cpu()
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref.abs().max()
    assert torch.equal(pred[:, 0][clear], ref.argmax(dim=1)[clear])


def test_binary_head_trains():
    """loss_by_feat with out_channels=1: gradients of the repeated filter sum into the single stored one."""
    K = 2
    _, m, sd = _pair(K, out_channels=1)
    sd1 = dict(sd)
    for name in ('decode_head.conv_seg', 'decode_head.aux_cls_seg'):
        sd1[name + '.weight'] = sd[name + '.weight'][:1].clone()
        sd1[name + '.bias'] = sd[name + '.bias'][:1].clone()
    m.load_state_dict(sd1)
    m.to(DEV).train()
    img, lab = synth.make_scene(2, 64, 64, K, seed=9, coarse=(4, 4))
    x = oracle.preprocess(img).to(DEV)
    samples = [dict(gt_sem_seg=dict(data=lab[i:i + 1].to(DEV))) for i in range(2)]
    total, log = m.parse_losses(m.loss(x, samples))
    total.backward()
    g = m.decode_head.conv_seg.weight.grad
    assert g is not None and g.shape == (1, 64, 1, 1) and torch.isfinite(g).all() and float(g.abs().max()) > 0
