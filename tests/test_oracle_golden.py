"""The CPU oracle against the golden vectors produced by the reference's OWN modules
(tests/golden/make_golden.py ran them verbatim in the build container)."""
import os

import numpy as np
import pytest
import torch

import oracle
import lednet_b200  # noqa: F401
from lednet_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.fixture(scope='module')
def k2_models():
    torch.manual_seed(0)
    bb = oracle.OracleLEDNet(3, 32, 128)
    hd = oracle.OracleLEDHead(128, 64, 2)
    bb.load_state_dict(synth.make_state_dict(bb.state_dict(), seed=2))
    hd.load_state_dict(synth.make_state_dict(hd.state_dict(), seed=3))
    return bb, hd


def test_param_counts_match_reference(golden_dir, k2_models):
    g = _load(golden_dir, 'r0_head_k2.npz')
    bb, hd = k2_models
    assert sum(p.numel() for p in bb.parameters()) == int(g['n_params_backbone']) == 5620576
    assert sum(p.numel() for p in hd.parameters()) == int(g['n_params_head']) == 112780


def test_r0_trunk_and_head_eval(golden_dir, k2_models):
    g = _load(golden_dir, 'r0_head_k2.npz')
    bb, hd = k2_models
    bb.eval(), hd.eval()
    x = oracle.preprocess(synth.make_images_u8(1, 64, 128, seed=0))
    with torch.no_grad():
        c5, x1, x2 = bb(x)
        xc, h1, h2 = hd((c5, x1, x2))
        fused = oracle.fuse_logits(xc, h1, h2)
        pred = oracle.postprocess_argmax(fused)
    for name, t in dict(c5=c5, x1=x1, x2=x2, xc=xc, h1=h1, h2=h2, fused=fused).items():
        # fp32 re-association across thread counts / ISAs: compare norm-wise
        err = np.abs(t.numpy() - g[name]).max() / np.abs(g[name]).max()
        assert err < 1e-5, (name, err)
    assert (pred[:, 0].numpy() == g['pred']).mean() > 0.9999


def test_fusion_odd_sizes(golden_dir):
    g = _load(golden_dir, 'fuse_odd.npz')
    out = oracle.fuse_logits(torch.from_numpy(g['xc']), torch.from_numpy(g['h1']),
                             torch.from_numpy(g['h2']))
    assert out.shape == (2, 2, 100, 52)
    np.testing.assert_array_equal(out.numpy(), g['fused'])


def test_iou_metric(golden_dir):
    g = _load(golden_dir, 'iou.npz')
    K = 19
    pred = torch.from_numpy(g['pred'].astype(np.int64))
    lab = torch.from_numpy(g['label'].astype(np.int64))
    res = [oracle.intersect_and_union(pred[i], lab[i], K, 255) for i in range(3)]
    areas = np.stack([torch.stack([r[j] for r in res]).numpy() for j in range(4)])
    np.testing.assert_array_equal(areas, g['areas'])
    # confusion matrix (rows GT, spill row K) reproduces the four histograms exactly
    for i in range(3):
        cm = oracle.confusion_matrix(g['pred'][i], g['label'][i], K, 255)
        for a, b in zip(oracle.confusion_to_areas(cm), res[i]):
            np.testing.assert_array_equal(np.asarray(a, dtype=np.float32), b.numpy())
    met = oracle.total_area_to_metrics(*[torch.from_numpy(g['areas'][j].sum(0)) for j in range(4)],
                                       ['mIoU', 'mDice', 'mFscore'])
    for k, v in met.items():
        np.testing.assert_array_equal(v, g['met_' + k])
    summ = oracle.compute_metrics(res)
    for k, v in summ.items():
        assert float(v) == float(g['sum_' + k]), k
    with pytest.raises(KeyError):
        oracle.total_area_to_metrics(*[torch.ones(K)] * 4, ['bogus'])


def test_ohem_and_accuracy(golden_dir):
    g = _load(golden_dir, 'ohem.npz')
    score = torch.from_numpy(g['score'])
    target = torch.from_numpy(g['target'].astype(np.int64))
    cw = g['class_weight'].tolist()
    for tag, kw in [('a', dict(thres=0.9, min_kept=500, loss_weight=1.0)),
                    ('b', dict(thres=0.3, min_kept=1200, loss_weight=0.4)),
                    ('c', dict(thres=0.7, min_kept=100000, loss_weight=1.0, class_weight=cw))]:
        s = score.clone().requires_grad_(True)
        loss = oracle.ohem_cross_entropy(s, target, **kw)
        loss.backward()
        np.testing.assert_allclose(loss.detach().numpy(), g['loss_' + tag], rtol=1e-6)
        np.testing.assert_allclose(s.grad.numpy(), g['grad_' + tag], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(oracle.accuracy(score, target, ignore_index=255).numpy(), g['acc'])
    # all-ignored target -> 0 (ohem_cross_entropy_loss.py:83-84)
    assert float(oracle.ohem_cross_entropy(score, torch.full_like(target, 255))) == 0.0


def test_train_losses(golden_dir, k2_models):
    g = _load(golden_dir, 'train_k2.npz')
    bb, hd = k2_models
    bb.train(), hd.train()
    hd.loss_decode = [dict(thres=0.9, min_kept=1000, loss_weight=1.0),
                      dict(thres=0.9, min_kept=1000, loss_weight=0.4)]
    x = oracle.preprocess(synth.make_images_u8(2, 64, 64, seed=9))
    lab = synth.make_labels(2, 64, 64, 2, seed=10, block=8)
    feats = bb(x)
    np.testing.assert_allclose(feats[0].detach().numpy(), g['c3'], rtol=1e-4, atol=1e-5)
    losses = hd.loss(feats, lab)
    for k in ('loss_context', 'loss_spatial', 'acc_seg'):
        np.testing.assert_allclose(losses[k].detach().numpy(), g[k], rtol=1e-4, err_msg=k)
    bb.eval(), hd.eval()


def test_sesp_block(golden_dir):
    """oracle/sesp.py against the reference's own SESP (eesp.py) outputs: bit-exact (same ATen ops)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(golden_dir, 'make_golden.py'))
    src = open(os.path.join(golden_dir, 'make_golden.py')).read()
    ns = {}
    # only the case table and the two seeded helpers are needed (the module's main() needs /root/reference)
    start, end = src.index('SESP_CASES = ['), src.index('def make_sesp')
    exec('import torch\nfrom lednet_b200 import synth\n' + src[start:end], ns)
    from oracle.sesp import OracleSESP
    g = _load(golden_dir, 'sesp.npz')
    for i, (tag, kw, shape) in enumerate(ns['SESP_CASES']):
        m = OracleSESP(**kw).eval()
        m.load_state_dict(ns['sesp_state_dict'](m.state_dict()))
        assert sum(p.numel() for p in m.parameters()) == int(g[tag + '_nparams'])
        with torch.no_grad():
            out = m(ns['sesp_input'](i, kw['nIn'], shape)).numpy()
        err = np.abs(out - g[tag]).max() / np.abs(g[tag]).max()
        assert err < 1e-6, (tag, err)


def test_mfaf_block(golden_dir):
    """oracle/mfaf.py against the reference's own Muti_AFF (classification/model_utils.py) outputs: bit-exact."""
    import block_cases as bc
    from oracle.mfaf import OracleMutiAFF
    g = _load(golden_dir, 'mfaf.npz')
    for i, (tag, kw, shape) in enumerate(bc.MFAF_CASES):
        m = OracleMutiAFF(**kw).eval()
        m.load_state_dict(bc.block_state_dict(m.state_dict(), seed=31))
        assert sum(p.numel() for p in m.parameters()) == int(g[tag + '_nparams'])
        x, r = bc.block_input(i, kw['channels'], shape, 300, n_inputs=2)
        with torch.no_grad():
            out = m(x, r).numpy()
        assert np.array_equal(out, g[tag]), tag


def test_getb_block(golden_dir):
    """oracle/getb.py against the reference's own GETBBlock (backbones/UNetFormer_GETB.py) outputs.  The oracle
    writes einops' rearranges as view/permute, so matmul operand strides may differ: 1e-6, not bit-exact."""
    import block_cases as bc
    from oracle.getb import OracleGETBBlock
    g = _load(golden_dir, 'getb.npz')
    for i, (tag, kw, shape) in enumerate(bc.GETB_CASES):
        m = OracleGETBBlock(**kw).eval()
        m.load_state_dict(bc.block_state_dict(m.state_dict(), seed=41))
        assert sum(p.numel() for p in m.parameters()) == int(g[tag + '_nparams'])
        with torch.no_grad():
            out = m(bc.block_input(i, kw['dim'], shape, 400)).numpy()
        assert out.shape == g[tag].shape
        assert np.abs(out - g[tag]).max() <= 1e-6 * np.abs(g[tag]).max(), tag


def test_seam_gate(golden_dir):
    """oracle/seam.py against the prototype's OWN statements (tools/speed/ddrnet_speed.py:282-338, 388-389, executed
    through an AST slice when the fixture was generated): edge response, 0/1 mask and gated output."""
    import block_cases as bc
    from oracle.seam import OracleSEAM
    g = np.load(os.path.join(golden_dir, 'seam.npz'))
    o = OracleSEAM(64).eval()
    o.load_state_dict(bc.seam_state_dict(o.state_dict()))
    for tag, shape in bc.SEAM_CASES:
        x, xs = bc.seam_inputs(shape)
        with torch.no_grad():
            mask, e = o.edge_mask(x)
            out = o(x, xs)
        # (conv arithmetic order depends on the host's thread count: compare to float rounding, masks outside the
        #  pixels whose Laplacian sits within rounding distance of the hard threshold)
        np.testing.assert_allclose(e.numpy(), g[tag + '_edge'], atol=2e-6, rtol=0)
        unstable, unstable3 = bc.seam_unstable(torch.from_numpy(g[tag + '_edge']))
        n = mask.numel()
        gm = torch.from_numpy(np.unpackbits(g[tag + '_mask'])[:n].reshape(mask.shape))
        assert not ((mask != gm) & ~unstable).any()
        assert unstable.float().mean() < 0.01
        ref = torch.from_numpy(g[tag])
        err = ((out[:, bc.SEAM_GOLDEN_CHANNELS] - ref).abs() * ~unstable3).max() / ref.abs().max()
        assert err < 2e-6, (tag, float(err))


def test_oracle_stack_batch_matches_reference_golden(golden_dir):
    """oracle.stack_batch + oracle.preprocess == the reference's stack_batch behind its normalisation statements
    (tests/golden/stack.npz, generated by make_golden.py from mmseg/utils/misc.py executed where it lies)."""
    import oracle
    from block_cases import STACK_CASES, stack_inputs
    gold = _load(golden_dir, 'stack.npz')
    for ci, (tag, shapes, size, div, pad_val, seg_pad_val) in enumerate(STACK_CASES):
        imgs, labs = stack_inputs(ci, shapes)
        xs = [oracle.preprocess(im[None])[0] for im in imgs]
        batch, lab, pads = oracle.stack_batch(xs, labs, size=size, size_divisor=div, pad_val=pad_val,
                                              seg_pad_val=seg_pad_val)
        np.testing.assert_array_equal(batch.numpy(), gold[tag + '_inputs'])
        np.testing.assert_array_equal(lab.numpy(), gold[tag + '_labels'].astype(np.int64))
        np.testing.assert_array_equal(np.array(pads), gold[tag + '_padding'])
