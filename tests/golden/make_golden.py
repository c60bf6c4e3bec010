"""Generate tests/golden/*.npz from the reference's OWN modules.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It executes the reference's files verbatim (oracle/ref_loader.py loads them by
path; nothing is copied) on seeded inputs and stores inputs' seeds + outputs.
Weights are NOT stored: they are regenerated from the seed by
``led-net_b200/synth.py`` (numpy PCG64, platform-stable), keeping fixtures small.

Fixtures
* r0_head_k2.npz  : DDRNet(channels=32,ppm=128) trunk with stem taps + LEDHead(K=2)
                    eval on 1x3x64x128 -> c5/x1/x2 features, 3 head logits, the
                    patched predict_by_feat fusion, argmax.
* fuse_odd.npz    : predict_by_feat on odd sizes (ceil paths), K=2.
* iou.npz         : IoUMetric.intersect_and_union + total_area_to_metrics + compute.
* ohem.npz        : OhemCrossEntropy (with and without class_weight) + accuracy,
                    forward values and d(loss)/d(score).
* train_k2.npz    : train-mode LEDHead.loss_by_feat losses on a 2x3x64x64 batch.
* sesp.npz        : SESP block (eesp.py) eval outputs for the three registered configurations
                    (64,64,Spatial) / (128,128,context,r_lim=9) / (256,256,context,r_lim=9) and a
                    channel-changing, non-V2 one; weights from synth seed 5 with PReLU slopes and BN
                    statistics made non-trivial (sesp_state_dict below).
* mfaf.npz        : Muti_AFF (classification/model_utils.py) eval outputs on tests/block_cases.MFAF_CASES.
* getb.npz        : GETBBlock (backbones/UNetFormer_GETB.py) eval outputs on tests/block_cases.GETB_CASES.
* stack.npz       : stack_batch (mmseg/utils/misc.py:30-128) behind SegDataPreProcessor's normalisation statements
                    (data_preprocessor.py:118-123) on tests/block_cases.STACK_CASES: padded batch, padded labels,
                    padding_size metainfo.
* led_trunk.npz   : the LED wiring = class DDRNet1 of the speed prototype, executed from its own file (AST) with the
                    reference's own blocks, eval mode, + stem taps: eight output channels, per-channel means of c5/x1/x2.
* seam.npz        : the SEAM edge gate of the authors' speed prototype (tools/speed/ddrnet_speed.py:282-338,388-389),
                    its own statements executed through oracle.ref_loader.load_seam (AST slice): the 0/1 edge mask
                    (bit-packed), the normalised edge response and the gated output on tests/block_cases.SEAM_CASES.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402
import lednet_b200  # noqa: E402,F401  (import shim for the hyphenated package dir)
from lednet_b200 import synth  # noqa: E402


class _PixelData:
    def __init__(self, data):
        self.data = data
        self.shape = data.shape[-2:]


class _Sample:
    def __init__(self, label):
        self.gt_sem_seg = _PixelData(label)


def ref_backbone_with_taps(ddr, x):
    """R0 = verbatim DDRNet forward plus the two stem taps (stem[0], stem[1])."""
    taps = {}
    h0 = ddr.stem[0].register_forward_hook(lambda m, i, o: taps.__setitem__('x1', o.clone()))
    h1 = ddr.stem[1].register_forward_hook(lambda m, i, o: taps.__setitem__('x2', o.clone()))
    out = ddr(x)
    h0.remove(), h1.remove()
    return out, taps['x1'], taps['x2']


SESP_CASES = [
    ('s64', dict(nIn=64, nOut=64, Spatial=True), (2, 11, 14)),
    ('c128', dict(nIn=128, nOut=128, Spatial=False, r_lim=9), (1, 12, 10)),
    ('c256', dict(nIn=256, nOut=256, Spatial=False, r_lim=9), (1, 7, 9)),
    ('x32_64_nov2', dict(nIn=32, nOut=64, Spatial=False, SESPV2=False), (1, 9, 8)),
]


def sesp_state_dict(template, seed=5):
    """synth weights + per-channel PReLU slopes in [0.05, 0.45] (make_state_dict gives a constant 0.25)."""
    sd = synth.make_state_dict(template, seed=seed)
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k.endswith('act.weight') or k == 'module_act.weight':
            sd[k] = 0.05 + 0.4 * torch.rand(sd[k].shape, generator=g)
    return sd


def sesp_input(case_index, nin, shape):
    g = torch.Generator().manual_seed(100 + case_index)
    return torch.randn(shape[0], nin, shape[1], shape[2], generator=g)


def make_sesp(ref):
    out = {}
    for i, (tag, kw, shape) in enumerate(SESP_CASES):
        m = ref.SESP(**kw).eval()
        m.load_state_dict(sesp_state_dict(m.state_dict()))
        with torch.no_grad():
            out[tag] = m(sesp_input(i, kw['nIn'], shape)).numpy()
        out[tag + '_nparams'] = np.int64(sum(p.numel() for p in m.parameters()))
    np.savez_compressed(os.path.join(OUT, 'sesp.npz'), **out)


def make_blocks(ref):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import block_cases as bc
    out = {}
    for i, (tag, kw, shape) in enumerate(bc.MFAF_CASES):
        m = ref.Muti_AFF(**kw).eval()
        m.load_state_dict(bc.block_state_dict(m.state_dict(), seed=31))
        x, r = bc.block_input(i, kw['channels'], shape, 300, n_inputs=2)
        with torch.no_grad():
            out[tag] = m(x, r).numpy()
        out[tag + '_nparams'] = np.int64(sum(p.numel() for p in m.parameters()))
    np.savez_compressed(os.path.join(OUT, 'mfaf.npz'), **out)
    out = {}
    for i, (tag, kw, shape) in enumerate(bc.GETB_CASES):
        m = ref.GETBBlock(**kw).eval()
        m.load_state_dict(bc.block_state_dict(m.state_dict(), seed=41))
        with torch.no_grad():
            out[tag] = m(bc.block_input(i, kw['dim'], shape, 400)).numpy()
        out[tag + '_nparams'] = np.int64(sum(p.numel() for p in m.parameters()))
    np.savez_compressed(os.path.join(OUT, 'getb.npz'), **out)


def make_seam():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import block_cases as bc
    m = ref_loader.load_seam()().eval()
    m.load_state_dict(bc.seam_state_dict(m.state_dict()))
    out = {}
    for tag, shape in bc.SEAM_CASES:
        x, xs = bc.seam_inputs(shape)
        with torch.no_grad():
            mask, e = m.edge(x)
            out[tag] = m(x, xs)[:, bc.SEAM_GOLDEN_CHANNELS].numpy()      # three of the 64 channels keep the fixture small
            out[tag + '_mask'] = np.packbits(mask.numpy().astype(np.uint8))
            out[tag + '_edge'] = e.numpy()
    np.savez_compressed(os.path.join(OUT, 'seam.npz'), **out)


class _StackSample:
    """The part of SegDataSample that stack_batch touches (misc.py:98-121)."""

    def __init__(self, label):
        self.gt_sem_seg = _PixelData(label)
        self.meta = {}

    def __contains__(self, key):
        return key == 'gt_sem_seg'

    def set_metainfo(self, d):
        self.meta.update(d)


class _LabelData:
    """PixelData stand-in whose `.data` can be deleted and re-assigned, `.shape` following it."""

    def __init__(self, data):
        self.data = data

    @property
    def shape(self):
        return tuple(self.data.shape[-2:])


def make_stack():
    """stack.npz: the reference's stack_batch on normalised float images + label maps (tests/block_cases.STACK_CASES);
    the normalisation in front of it is data_preprocessor.py:118-123's three statements, applied here verbatim."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import block_cases as bc
    stack_batch = ref_loader.load_stack_batch()
    mean = torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1)
    std = torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1)
    out = {}
    for ci, (tag, shapes, size, div, pad_val, seg_pad_val) in enumerate(bc.STACK_CASES):
        imgs, labs = bc.stack_inputs(ci, shapes)
        inputs = [_input[[2, 1, 0], ...] for _input in imgs]
        inputs = [_input.float() for _input in inputs]
        inputs = [(_input - mean) / std for _input in inputs]
        samples = []
        for lab in labs:
            smp = _StackSample(lab)
            smp.gt_sem_seg = _LabelData(lab.clone())
            samples.append(smp)
        batch, samples = stack_batch(inputs=inputs, data_samples=samples, size=size, size_divisor=div,
                                     pad_val=pad_val, seg_pad_val=seg_pad_val)
        out[tag + '_inputs'] = batch.numpy()
        out[tag + '_labels'] = torch.stack([smp.gt_sem_seg.data for smp in samples]).numpy().astype(np.int16)
        out[tag + '_padding'] = np.array([smp.meta['padding_size'] for smp in samples], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, 'stack.npz'), **out)


def make_led():
    """led_trunk.npz: the prototype's DDRNet1 (tools/speed/ddrnet_speed.py:39-406, executed through
    oracle.ref_loader.load_ddrnet1) in eval mode + the two stem taps, on tests/block_cases.LED_CASES."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import block_cases as bc
    m = ref_loader.load_ddrnet1()().eval()
    m.load_state_dict(bc.led_state_dict(m.state_dict()))
    out = {}
    for ci, (tag, shape) in enumerate(bc.LED_CASES):
        x = bc.led_input(ci, shape)
        with torch.no_grad():
            c5, x1, x2 = m.forward_with_taps(x)
        out[tag + '_c5'] = c5[:, bc.LED_GOLDEN_CHANNELS].numpy()
        out[tag + '_c5_absmax'] = np.array(float(c5.abs().max()))
        out[tag + '_c5_chan_mean'] = c5.mean(dim=(0, 2, 3)).numpy()          # every channel takes part in the check
        out[tag + '_x1_chan_mean'] = x1.mean(dim=(0, 2, 3)).numpy()
        out[tag + '_x2_chan_mean'] = x2.mean(dim=(0, 2, 3)).numpy()
    np.savez_compressed(os.path.join(OUT, 'led_trunk.npz'), **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    if 'led' in sys.argv[1:]:             # regenerate only the LED-wiring fixture
        make_led()
        return
    if 'stack' in sys.argv[1:]:           # regenerate only the stack_batch fixture
        make_stack()
        return
    if 'seam' in sys.argv[1:]:            # regenerate only the SEAM fixture
        make_seam()
        return
    ref = ref_loader.load()
    if 'sesp' in sys.argv[1:]:            # regenerate only the SESP fixture
        make_sesp(ref)
        return
    if 'blocks' in sys.argv[1:]:          # regenerate only the MFAF / GETB fixtures
        make_blocks(ref)
        return
    make_sesp(ref)
    make_blocks(ref)
    make_seam()
    make_stack()
    make_led()

    # ---------------- r0_head_k2 -------------------------------------------------
    ddr = ref.DDRNet(in_channels=3, channels=32, ppm_channels=128,
                     norm_cfg=dict(type='BN', requires_grad=True), align_corners=False)
    head = ref.LEDHead(in_channels=128, channels=64, num_classes=2, dropout_ratio=0.,
                       norm_cfg=dict(type='BN', requires_grad=True), align_corners=False,
                       loss_decode=[dict(type='OhemCrossEntropy', thres=0.9, min_kept=131072,
                                         loss_weight=1.0),
                                    dict(type='OhemCrossEntropy', thres=0.9, min_kept=131072,
                                         loss_weight=0.4)])
    ddr.load_state_dict(synth.make_state_dict(ddr.state_dict(), seed=2))
    head.load_state_dict(synth.make_state_dict(head.state_dict(), seed=3))
    ddr.eval(), head.eval()
    from oracle import preprocess
    img = synth.make_images_u8(1, 64, 128, seed=0)
    x = preprocess(img)
    with torch.no_grad():
        c5, x1, x2 = ref_backbone_with_taps(ddr, x)
        xc, h1, h2 = head.forward((c5, x1, x2))
        fused = head.predict_by_feat((xc, h1, h2), [dict(img_shape=(64, 128))])
        pred = fused.argmax(dim=1)
    np.savez_compressed(os.path.join(OUT, 'r0_head_k2.npz'), c5=c5.numpy(), x1=x1.numpy(),
                        x2=x2.numpy(), xc=xc.numpy(), h1=h1.numpy(), h2=h2.numpy(),
                        fused=fused.numpy(), pred=pred.numpy().astype(np.uint8),
                        n_params_backbone=sum(p.numel() for p in ddr.parameters()),
                        n_params_head=sum(p.numel() for p in head.parameters()))

    # ---------------- fuse_odd ---------------------------------------------------
    g = np.random.default_rng(11)
    xc_o = torch.from_numpy(g.normal(size=(2, 2, 13, 7)).astype(np.float32))
    h2_o = torch.from_numpy(np.maximum(g.normal(size=(2, 2, 25, 13)), 0).astype(np.float32))
    h1_o = torch.from_numpy(np.maximum(g.normal(size=(2, 2, 50, 26)), 0).astype(np.float32))
    with torch.no_grad():
        fo = head.predict_by_feat((xc_o, h1_o, h2_o), [dict(img_shape=(100, 52))])
    np.savez_compressed(os.path.join(OUT, 'fuse_odd.npz'), xc=xc_o.numpy(), h1=h1_o.numpy(),
                        h2=h2_o.numpy(), fused=fo.numpy())

    # ---------------- iou --------------------------------------------------------
    K = 19
    pred_i = torch.from_numpy(g.integers(0, K, (3, 96, 160)))
    lab_i = synth.make_labels(3, 96, 160, K, seed=5)
    lab_i[0, :4, :4] = 40              # out-of-range, not ignore: histc drops it from area_label
    res = [ref.IoUMetric.intersect_and_union(pred_i[i], lab_i[i], K, 255) for i in range(3)]
    cols = tuple(zip(*res))
    tot = [sum(c) for c in cols]
    met = ref.IoUMetric.total_area_to_metrics(*tot, ['mIoU', 'mDice', 'mFscore'], None, 1)
    m = ref.IoUMetric(ignore_index=255, iou_metrics=['mIoU'])
    m.dataset_meta = dict(classes=[str(i) for i in range(K)])
    summary = m.compute_metrics(list(res))
    np.savez_compressed(
        os.path.join(OUT, 'iou.npz'), pred=pred_i.numpy().astype(np.uint8),
        label=lab_i.numpy().astype(np.uint8),
        areas=np.stack([torch.stack(list(c)).numpy() for c in cols]),   # [4,3,K]
        **{'met_' + k: v for k, v in met.items()},
        **{'sum_' + k: np.float64(v) for k, v in summary.items()})

    # ---------------- ohem -------------------------------------------------------
    Kc = 5
    score = torch.from_numpy(g.normal(scale=2.0, size=(2, Kc, 24, 40)).astype(np.float32))
    target = synth.make_labels(2, 24, 40, Kc, seed=7, block=8)
    cw = [0.8, 1.2, 1.0, 0.5, 1.5]
    out = dict(score=score.numpy(), target=target.numpy().astype(np.uint8), class_weight=np.array(cw))
    for tag, kw in [('a', dict(thres=0.9, min_kept=500, loss_weight=1.0)),
                    ('b', dict(thres=0.3, min_kept=1200, loss_weight=0.4)),
                    ('c', dict(thres=0.7, min_kept=100000, loss_weight=1.0, class_weight=cw))]:
        s = score.clone().requires_grad_(True)
        loss = ref.OhemCrossEntropy(**kw)(s, target)
        loss.backward()
        out['loss_' + tag] = loss.detach().numpy()
        out['grad_' + tag] = s.grad.numpy()
    out['acc'] = ref.accuracy(score, target, ignore_index=255).numpy()
    np.savez_compressed(os.path.join(OUT, 'ohem.npz'), **out)

    # ---------------- train_k2 ---------------------------------------------------
    ddr.train(), head.train()
    img_t = synth.make_images_u8(2, 64, 64, seed=9)
    lab_t = synth.make_labels(2, 64, 64, 2, seed=10, block=8)
    xt = preprocess(img_t)
    (c3, c5t), x1t, x2t = ref_backbone_with_taps(ddr, xt)
    logits = head.forward((c3, c5t, x1t, x2t))
    samples = [_Sample(lab_t[i:i + 1]) for i in range(2)]
    # small min_kept so the k-th statistic (not only `thres`) is exercised
    for l in head.loss_decode:
        l.min_kept = 1000
    losses = head.loss_by_feat(logits, samples)
    np.savez_compressed(os.path.join(OUT, 'train_k2.npz'),
                        **{k: v.detach().numpy() for k, v in losses.items()},
                        c3=c3.detach().numpy())
    for f in sorted(os.listdir(OUT)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
