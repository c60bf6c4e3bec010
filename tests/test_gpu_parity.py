"""GPU parity tests proper: every kernel and the whole path, through the C ABI, against the CPU
oracle (which tests/test_oracle_golden.py pins to the reference's own modules)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
import lednet_b200 as L
from lednet_b200 import ops, synth
from util import build_pair, rel_err, near_tie_mask, bf16_storage_agreement

pytestmark = pytest.mark.gpu
DEV = 'cuda'


# ------------------------------------------------------------------ convolution (CUDA-core path)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('cin,cout,k,stride,hw', [
    (3, 32, 3, 2, (34, 50)), (32, 32, 3, 1, (20, 36)), (32, 64, 3, 2, (24, 40)), (64, 19, 1, 1, (9, 13)),
    (64, 128, 1, 2, (16, 24)), (32, 2, 3, 1, (17, 33)), (128, 64, 3, 1, (8, 16))])
def test_conv_direct(dtype, cin, cout, k, stride, hw):
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(2, cin, *hw, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ps, pb = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.1
    xq = x.to(dtype).float()
    for use_pre, use_res, relu in [(False, False, False), (True, True, True)]:
        xin = F.relu(xq * ps.view(1, -1, 1, 1) + pb.view(1, -1, 1, 1)) if use_pre else xq
        wq = w.to(dtype).float() if False else w
        ref = F.conv2d(xin, wq, b, stride, k // 2)
        res = torch.randn(ref.shape, generator=g).to(dtype).float() if use_res else None
        if use_res:
            ref = ref + res
        if relu:
            ref = F.relu(ref)
        out = ops.conv2d(xq.permute(0, 2, 3, 1).contiguous().to(DEV, dtype), w, b, stride, relu,
                         None if res is None else res.permute(0, 2, 3, 1).contiguous().to(DEV, dtype),
                         ps if use_pre else None, pb if use_pre else None, backend=1)
        got = out.float().cpu().permute(0, 3, 1, 2)
        tol = 1e-5 if dtype == torch.float32 else 6e-3    # bf16: output rounding only (inputs pre-rounded)
        assert rel_err(got, ref) < tol, (use_pre, rel_err(got, ref))


# ------------------------------------------------------------------ fused tail
def _tail_inputs(n, k, hc, wc, h4, w4, h2, w2, seed):
    g = np.random.default_rng(seed)
    xc = torch.from_numpy(g.normal(scale=2.0, size=(n, k, hc, wc)).astype(np.float32))
    hx2 = torch.from_numpy(np.maximum(g.normal(size=(n, k, h4, w4)), 0).astype(np.float32))
    hx1 = torch.from_numpy(np.maximum(g.normal(size=(n, k, h2, w2)), 0).astype(np.float32))
    return xc, hx2, hx1


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('k,sizes', [
    (19, (8, 16, 16, 32, 32, 64)), (2, (8, 16, 16, 32, 32, 64)), (19, (13, 7, 25, 13, 50, 26)),
    (3, (1, 1, 2, 2, 4, 4)), (19, (5, 40, 10, 80, 20, 160)), (150, (4, 4, 8, 8, 16, 16))])
def test_tail_fuse_argmax(dtype, k, sizes):
    hc, wc, h4, w4, h2, w2 = sizes
    xc, hx2, hx1 = _tail_inputs(2, k, hc, wc, h4, w4, h2, w2, seed=k + hc)
    xc, hx2, hx1 = (t.to(dtype).float() for t in (xc, hx2, hx1))
    ref = oracle.fuse_logits(xc, hx1, hx2)
    ref_pred = oracle.postprocess_argmax(ref)[:, 0]
    for pd in (torch.uint8, torch.int64):
        pred, logits = ops.head_fuse_argmax(xc.to(DEV, dtype), hx2.to(DEV, dtype), hx1.to(DEV, dtype),
                                            pred_dtype=pd, want_logits=True)
        assert pred.dtype == pd and tuple(pred.shape) == (2, 2 * h2, 2 * w2)
        assert rel_err(logits.cpu(), ref) < 2e-6
        mism = pred.cpu().long() != ref_pred
        # any disagreement must sit on a numerical tie of the two best logits
        assert not (mism & ~near_tie_mask(ref, 1e-6)).any()
        assert mism.float().mean() < 1e-4
        # and the kernel's own argmax is exactly the first-max argmax of its own logits
        assert torch.equal(pred.long(), logits.argmax(1))
    # without logits: same labels
    pred2, none = ops.head_fuse_argmax(xc.to(DEV, dtype), hx2.to(DEV, dtype), hx1.to(DEV, dtype))
    assert none is None and torch.equal(pred2.long(), pred.long())


def test_tail_first_max_tie_break():
    # all-equal logits: argmax must return class 0 everywhere (torch.argmax: first maximal index)
    z = torch.zeros(1, 19, 4, 4, device=DEV)
    pred, _ = ops.head_fuse_argmax(z, torch.zeros(1, 19, 8, 8, device=DEV), torch.zeros(1, 19, 16, 16, device=DEV))
    assert int(pred.max()) == 0
    # golden: the reference's own predict_by_feat on odd sizes
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fuse_odd.npz'))
    pred, logits = ops.head_fuse_argmax(*(torch.from_numpy(g[n]).to(DEV) for n in ('xc', 'h2', 'h1')), want_logits=True)
    assert rel_err(logits.cpu(), torch.from_numpy(g['fused'])) < 2e-6


def test_tail_rejects_bad_ladder():
    with pytest.raises(L.LedB200Error):
        ops.head_fuse_argmax(torch.zeros(1, 2, 64, 64, device=DEV), torch.zeros(1, 2, 8, 8, device=DEV),
                             torch.zeros(1, 2, 16, 16, device=DEV))


# ------------------------------------------------------------------ confusion matrix / IoU
@pytest.mark.parametrize('pd,gd', [(torch.uint8, torch.uint8), (torch.int64, torch.int64),
                                   (torch.uint8, torch.int64), (torch.int64, torch.uint8)])
@pytest.mark.parametrize('k,n', [(19, 96 * 160 * 3), (2, 1000003), (150, 4099), (19, 5), (19, 0)])
def test_confusion_matrix_bit_exact(pd, gd, k, n):
    g = np.random.default_rng(k + n)
    pred = g.integers(0, k, n).astype(np.int64)
    gt = g.integers(0, k, n).astype(np.int64)
    gt[g.random(n) < 0.07] = 255
    if n > 10:
        gt[:3] = 200 if k < 200 else 254          # out-of-range, not ignore -> spill row
    cm = ops.confusion_accumulate(torch.from_numpy(pred).to(DEV, pd), torch.from_numpy(gt).to(DEV, gd), k, 255)
    ref = oracle.confusion_matrix(pred, gt, k, 255)
    np.testing.assert_array_equal(cm.cpu().numpy(), ref)
    # accumulation (not overwrite)
    cm2 = ops.confusion_accumulate(torch.from_numpy(pred).to(DEV, pd), torch.from_numpy(gt).to(DEV, gd), k, 255, cm)
    np.testing.assert_array_equal(cm2.cpu().numpy(), 2 * ref)


def test_iou_metric_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'iou.npz'))
    K = 19
    m = L.IoUMetric(iou_metrics=['mIoU'])
    m.dataset_meta = dict(classes=[str(i) for i in range(K)])
    pred = torch.from_numpy(g['pred']).to(DEV)
    lab = torch.from_numpy(g['label']).to(DEV)
    for i in range(3):
        got = L.IoUMetric.intersect_and_union(pred[i].long(), lab[i].long(), K, 255)
        for j in range(4):
            np.testing.assert_array_equal(got[j].numpy(), g['areas'][j][i])     # bit-exact vs histc
    m.process(None, [{'pred_sem_seg': {'data': pred[i][None]}, 'gt_sem_seg': {'data': lab[i][None]}}
                     for i in range(3)])
    s = m.compute_metrics()
    for k, v in s.items():
        assert float(v) == float(g['sum_' + k]), k
    # batched fast path gives the same totals
    m2 = L.IoUMetric()
    m2.dataset_meta = m.dataset_meta
    m2.process_batch(pred, lab)
    assert torch.equal(m2.total_confusion(), m.total_confusion())


def test_confusion_full_size_properties():
    # BASELINE config 2 size: 16 x 1024 x 2048, K=19 - size-independent properties
    K, n = 19, 16 * 1024 * 2048
    g = torch.Generator(device=DEV).manual_seed(0)
    pred = torch.randint(0, K, (n,), device=DEV, generator=g, dtype=torch.int64).to(torch.uint8)
    gt = torch.randint(0, K, (n,), device=DEV, generator=g, dtype=torch.int64).to(torch.uint8)
    gt[torch.rand(n, device=DEV, generator=g) < 0.05] = 255
    cm = ops.confusion_accumulate(pred, gt, K)
    assert int(cm.sum()) == int((gt != 255).sum())                      # every kept pixel counted once
    assert torch.equal(cm[:K].sum(1), torch.bincount(gt[gt != 255].long(), minlength=K))
    assert torch.equal(cm.sum(0), torch.bincount(pred[gt != 255].long(), minlength=K))
    # linearity: halves add up
    h = n // 2
    a = ops.confusion_accumulate(pred[:h], gt[:h], K)
    b = ops.confusion_accumulate(pred[h:], gt[h:], K)
    assert torch.equal(a + b, cm)
    # pred == gt -> diagonal only
    d = ops.confusion_accumulate(gt.clamp(max=K - 1), gt, K)
    assert int(d.sum()) == int(torch.diagonal(d[:K]).sum())


# ------------------------------------------------------------------ OHEM CE
def test_ohem_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'ohem.npz'))
    score = torch.from_numpy(g['score']).to(DEV)
    target = torch.from_numpy(g['target'].astype(np.int64)).to(DEV)
    cw = g['class_weight'].tolist()
    for tag, kw in [('a', dict(thres=0.9, min_kept=500, loss_weight=1.0)),
                    ('b', dict(thres=0.3, min_kept=1200, loss_weight=0.4)),
                    ('c', dict(thres=0.7, min_kept=100000, loss_weight=1.0, class_weight=cw))]:
        s = score.clone().requires_grad_(True)
        loss = L.OhemCrossEntropy(**kw)(s, target)
        loss.backward()
        np.testing.assert_allclose(loss.item(), g['loss_' + tag], rtol=1e-5)     # north_star: 1e-2
        ge, gr = s.grad.cpu().numpy(), g['grad_' + tag]
        assert np.abs(ge - gr).max() <= 1e-5 * np.abs(gr).max() + 1e-9
    acc = L.accuracy(score, target, ignore_index=255)
    np.testing.assert_allclose(acc.cpu().numpy(), g['acc'], rtol=1e-6)
    # all pixels ignored -> 0 (ohem_cross_entropy_loss.py:83-84)
    assert float(L.OhemCrossEntropy()(score, torch.full_like(target, 255))) == 0.0


def test_ohem_vs_oracle_larger():
    K = 19
    g = np.random.default_rng(3)
    score = torch.from_numpy(g.normal(scale=3.0, size=(2, K, 128, 256)).astype(np.float32))
    target = synth.make_labels(2, 128, 256, K, seed=4)
    for kw in (dict(thres=0.9, min_kept=131072), dict(thres=0.05, min_kept=20000), dict(thres=0.7, min_kept=1)):
        ref = oracle.ohem_cross_entropy(score, target, **kw)
        got = L.OhemCrossEntropy(**kw)(score.to(DEV), target.to(DEV))
        np.testing.assert_allclose(got.item(), ref.item(), rtol=2e-5)


# ------------------------------------------------------------------ whole path
@pytest.mark.parametrize('k,hw', [(2, (64, 128)), (19, (96, 160)), (19, (72, 104))])
def test_full_path_fp32(k, hw):
    o, m = build_pair(k, dtype='fp32')
    img = synth.make_images_u8(2, *hw, seed=hw[0])
    x = oracle.preprocess(img)
    ref_logits, ref_pred = o.predict(x)
    pred, logits = m.engine().forward_infer(x.to(DEV), want_logits=True)
    assert tuple(logits.shape) == tuple(ref_logits.shape)
    err = rel_err(logits.cpu(), ref_logits)
    assert err < 1e-4, err                                               # north_star fp32 gate
    mism = pred.cpu().long() != ref_pred[:, 0]
    assert not (mism & ~near_tie_mask(ref_logits, 1e-4)).any()
    assert mism.float().mean() < 1e-3
    # raw uint8 BGR input with the preprocessing fused into the stem == normalised float input
    pred_u8 = m.predict_labels(img.to(DEV))
    assert (pred_u8 != pred).float().mean() < 1e-3
    # int64 predictions, reference-shaped predict()
    res = m.predict(x.to(DEV))
    assert len(res) == 2 and res[0]['pred_sem_seg']['data'].dtype == torch.int64
    assert tuple(res[0]['seg_logits']['data'].shape) == (k,) + tuple(ref_logits.shape[2:])


def test_full_path_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'r0_head_k2.npz'))
    o, m = build_pair(2, dtype='fp32')
    # the golden used seeds 2 (backbone) / 3 (head)
    sd = synth.make_state_dict(o.backbone.state_dict(), seed=2)
    hd = synth.make_state_dict(o.decode_head.state_dict(), seed=3)
    full = {'backbone.' + a: b for a, b in sd.items()}
    full.update({'decode_head.' + a: b for a, b in hd.items()})
    m.load_state_dict(full)
    x = oracle.preprocess(synth.make_images_u8(1, 64, 128, seed=0)).to(DEV)
    pred, logits = m.engine().forward_infer(x, want_logits=True)
    assert rel_err(logits.cpu(), torch.from_numpy(g['fused'])) < 1e-4
    assert (pred.cpu().numpy() == g['pred']).mean() > 0.999
    # module-level API: LEDNet.forward / LEDHead.forward contracts
    m.backbone.load_state_dict(sd), m.decode_head.load_state_dict(hd)
    m.backbone.set_compute_dtype('fp32'), m.decode_head.set_compute_dtype('fp32')
    c5, x1, x2 = m.backbone(x)
    for name, t in (('c5', c5), ('x1', x1), ('x2', x2)):
        assert rel_err(t.cpu(), torch.from_numpy(g[name])) < 1e-4, name
    xc, h1, h2 = m.decode_head((c5, x1, x2))
    for name, t in (('xc', xc), ('h1', h1), ('h2', h2)):
        assert rel_err(t.cpu(), torch.from_numpy(g[name])) < 1e-4, name
    fused = m.decode_head.predict_by_feat((xc, h1, h2))
    assert rel_err(fused.cpu(), torch.from_numpy(g['fused'])) < 1e-4


def test_layerwise_fp32():
    """Localises a mismatch: internal activations against oracle forward hooks."""
    o, m = build_pair(19, dtype='fp32')
    x = oracle.preprocess(synth.make_images_u8(1, 64, 128, seed=5))
    feats = {}
    hooks = []
    bb = o.backbone
    taps = {'x1': bb.stem[0], 'x2': bb.stem[1], 'backbone.stem.2.0': bb.stem[2][0],
            'backbone.context_branch_layers.0.1': bb.context_branch_layers[0][1],
            'backbone.spatial_branch_layers.2.0': bb.spatial_branch_layers[2][0],
            'backbone.context_branch_layers.2.0': bb.context_branch_layers[2][0],
            'spp.out': bb.spp}
    for n, mod in taps.items():
        hooks.append(mod.register_forward_hook(lambda mod, i, out, n=n: feats.__setitem__(n, out.detach().clone())))
    with torch.no_grad():
        o.predict(x)
    for h in hooks:
        h.remove()
    m.engine().forward_infer(x.to(DEV))
    for n, ref in feats.items():
        got = m.engine().debug_fetch(n)
        assert tuple(got.shape) == tuple(ref.shape), n
        assert rel_err(got, ref) < 1e-4, (n, rel_err(got, ref))


@pytest.mark.parametrize('k,hw', [(19, (128, 256)), (2, (64, 64))])
def test_full_path_bf16(k, hw):
    o, m = build_pair(k, dtype='bf16')
    img = synth.make_images_u8(2, *hw, seed=7)
    x = oracle.preprocess(img)
    ref_logits, ref_pred = o.predict(x)
    pred, logits = m.engine().forward_infer(x.to(DEV), want_logits=True)
    err = rel_err(logits.cpu(), ref_logits)
    agree = (pred.cpu().long() == ref_pred[:, 0]).float().mean().item()
    assert err < 2e-2, err                                               # north_star bf16 logit gate
    # argmax: the contractual gate (north_star: >= 99.9 % with trained margins) is tests/test_gpu_fullsize.py, at the
    # benchmarked sizes with trained weights.  THIS test runs random-init weights on a few thousand pixels, where the
    # oracle ITSELF, run with bf16 storage, agrees with its fp32 run on only ~99.5 % of pixels (`ceiling`, measured
    # below): a diagnostic band around that ceiling (0.5 %: 40 pixels of the 64 x 64 case, whose count moves by +-10
    # whenever a kernel changes where it rounds), >= 99 % absolute, and every flipped pixel a near-tie inside the bf16
    # logit tolerance (|top1 - top2| < 2 * 2e-2 * max|logit|), i.e. no flip that the tolerance does not explain.
    ceiling = bf16_storage_agreement(o, x, ref_pred)
    assert agree >= ceiling - 5e-3, (agree, ceiling)
    assert agree >= 0.99, agree
    mism = pred.cpu().long() != ref_pred[:, 0]
    assert not (mism & ~near_tie_mask(ref_logits, 4e-2)).any()
    # confusion matrix: bit-exact given identical predictions (oracle histc vs kernel int64)
    lab = synth.make_labels(2, *ref_pred.shape[-2:], k, seed=8)
    cm = ops.confusion_accumulate(pred, lab.to(DEV), k)
    for i_ref, i_got in zip(oracle.confusion_to_areas(oracle.confusion_matrix(pred.cpu().numpy(), lab.numpy(), k, 255)),
                            oracle.confusion_to_areas(cm.cpu().numpy())):
        np.testing.assert_array_equal(i_ref, i_got)


def test_slide_inference_matches_oracle():
    o, m = build_pair(2, dtype='fp32')
    o.test_cfg = dict(mode='slide', crop_size=(64, 64), stride=(48, 48))
    m.test_cfg = dict(o.test_cfg)
    x = oracle.preprocess(synth.make_images_u8(1, 96, 128, seed=11))
    ref = o.inference(x)
    got = m.inference(x.to(DEV))
    assert rel_err(got.cpu(), ref) < 1e-4


def _ref_postprocess(lg, padding, flip, ori_shape, align_corners, threshold=0.3):
    """the reference's postprocess_result body for one image (base.py:153-198) on CPU tensors"""
    K, H, W = lg.shape
    l, r, t, b = padding
    x = lg[None, :, t:H - b, l:W - r]
    if flip == 'horizontal':
        x = x.flip(dims=(3,))
    elif flip == 'vertical':
        x = x.flip(dims=(2,))
    x = oracle.resize(x, ori_shape, align_corners).squeeze(0)
    if K > 1:
        return x, x.argmax(dim=0, keepdim=True)
    x = x.sigmoid()
    return x, (x > threshold).to(x)


@pytest.mark.parametrize('K,hw,padding,flip,ori,ac', [
    (2, (64, 114), (0, 0, 0, 0), None, (90, 160), False),        # Apple-Branch ratio 1.40625 (512x910 -> 720x1280)
    (19, (40, 56), (0, 3, 0, 5), 'horizontal', (64, 96), False),  # padded + flipped + up
    (5, (37, 29), (2, 1, 3, 0), 'vertical', (20, 17), True),      # down-sampling, align_corners
    (3, (16, 24), (0, 0, 0, 0), None, (16, 24), False),           # identity
    (1, (20, 30), (0, 2, 0, 2), None, (31, 47), False),           # single class: sigmoid > threshold
])
def test_postprocess_result_matches_reference_steps(K, hw, padding, flip, ori, ac):
    g = torch.Generator().manual_seed(K * 100 + hw[0])
    lg = torch.randn(K, *hw, generator=g) * 3
    ref_lg, ref_pred = _ref_postprocess(lg, padding, flip, ori, ac)
    pred, out = L.ops.postprocess(lg.to(DEV), padding=padding, flip=flip, ori_shape=ori, align_corners=ac)
    assert out.shape == ref_lg.shape and pred.shape == ref_pred.shape
    assert rel_err(out.cpu(), ref_lg) < 1e-5
    if K > 1:
        assert pred.dtype == torch.int64
        mism = pred.cpu() != ref_pred
        # labels identical except numerical ties of the two best logits
        top2 = ref_lg.topk(2, dim=0).values
        assert not (mism[0] & ((top2[0] - top2[1]) > 1e-5 * ref_lg.abs().max())).any()
        assert mism.float().mean() < 1e-3
        assert torch.equal(out.argmax(dim=0, keepdim=True).cpu(), pred.cpu())       # self-consistent
    else:
        near = (ref_lg - 0.3).abs() < 1e-6
        assert torch.equal(pred.cpu()[~near], ref_pred[~near])


def test_predict_with_metainfo_uses_postprocess():
    """EncoderDecoder.predict with ori_shape / padding / flip in the samples' metainfo == oracle whole inference followed
    by the reference's post-processing steps."""
    K = 3
    o, m = build_pair(K, dtype='fp32')
    img = synth.make_images_u8(2, 64, 96, seed=5)
    x = oracle.preprocess(img)
    ref_logits, _ = o.predict(x)
    metas = [dict(metainfo=dict(ori_shape=(80, 100), img_padding_size=(0, 4, 0, 8))),
             dict(metainfo=dict(ori_shape=(64, 96), flip=True, flip_direction='horizontal'))]
    res = m.predict(x.to(DEV), metas)
    for i, meta in enumerate(metas):
        mi = meta['metainfo']
        ref_lg, ref_pred = _ref_postprocess(ref_logits[i], mi.get('img_padding_size', (0, 0, 0, 0)),
                                            mi.get('flip_direction') if mi.get('flip') else None, mi['ori_shape'], False)
        got = res[i]['seg_logits']['data'].cpu()
        assert got.shape == ref_lg.shape
        assert rel_err(got, ref_lg) < 1e-4
        assert (res[i]['pred_sem_seg']['data'].cpu() != ref_pred).float().mean() < 1e-3
