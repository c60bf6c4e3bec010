"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the registry
surface mirrors the reference's, host logic (metrics math, synthetic data) is right, and the
product never routes through the oracle or a CPU fallback."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
import torch

import oracle
import lednet_b200 as L
from lednet_b200 import lib as libmod, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'ledb200.h')).read()
    declared = set(re.findall(r'\b(ledb200_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'ledb200_cfg', 'ledb200_handle'}
    assert declared == set(libmod.SYMBOLS), declared ^ set(libmod.SYMBOLS)
    lib = libmod.get()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.ledb200_version() == 100


def test_abi_struct_matches_header():
    # 10 int32 + 6 float + 1 int32 + 7 reserved int32
    assert ctypes.sizeof(libmod.Cfg) == 4 * (10 + 6 + 1 + 7)


def test_no_cpu_fallback_and_no_oracle_import():
    pkg = os.path.join(ROOT, 'led-net_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            src = open(os.path.join(pkg, f)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, f
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64,
                                                       num_classes=2, dropout_ratio=0.)).eval()
    if not torch.cuda.is_available():
        with pytest.raises(L.LedB200Error):
            m.predict_labels(torch.zeros(1, 3, 64, 64))
        with pytest.raises(L.LedB200Error):
            L.ops.confusion_accumulate(torch.zeros(8, dtype=torch.uint8), torch.zeros(8, dtype=torch.uint8), 2)


def test_registry_surface_and_errors():
    for name in ('LEDNet', 'LEDHead', 'OhemCrossEntropy', 'EncoderDecoder', 'SegDataPreProcessor'):
        assert name in L.MODELS
    assert 'IoUMetric' in L.METRICS
    with pytest.raises(KeyError):
        L.MODELS.build(dict(type='NoSuchModel'))
    with pytest.raises(ValueError):      # decode_head.py:128-133
        L.LEDHead(128, 64, num_classes=19, out_channels=3, dropout_ratio=0.)
    with pytest.raises(TypeError):       # decode_head.py:149-151
        L.LEDHead(128, 64, num_classes=19, dropout_ratio=0., loss_decode='ce')
    with pytest.warns(UserWarning):      # decode_head.py:120-125
        h = L.LEDHead(128, 64, num_classes=2, dropout_ratio=0.)
    assert h.loss_decode[0].loss_name == 'loss_ohem' and h.loss_decode[1].loss_weight == 0.4
    loss = L.MODELS.build(dict(type='OhemCrossEntropy', thres=0.9, min_kept=0))
    assert loss.min_kept == 1           # ohem_cross_entropy_loss.py:47
    m = L.METRICS.build(dict(type='IoUMetric', iou_metrics=['mIoU']))
    with pytest.raises(KeyError):        # iou_metric.py:250-251
        m.total_area_to_metrics(*[torch.ones(3)] * 4, ['bogus'])


def test_state_dict_names_match_oracle_and_reference_layout():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64,
                                                       num_classes=19, dropout_ratio=0.))
    o = oracle.OracleSegmentor(num_classes=19)
    a = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert a == b
    for k in ('backbone.stem.0.conv.weight', 'backbone.stem.2.0.conv1.bn.running_var',
              'backbone.stem.4.0.downsample.0.weight', 'backbone.spp.scales.1.1.bn.weight',
              'decode_head.head.0.bn.weight', 'decode_head.head.0.conv.weight', 'decode_head.head.1.weight',
              'decode_head.conv_seg.bias', 'decode_head.aux_cls_seg.weight'):
        assert k in a, k
    # loading a reference-layout checkpoint works and invalidates the engine
    m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=4))
    assert m._engine is None


def test_metrics_host_math_matches_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, 'iou.npz'))
    tot = [torch.from_numpy(g['areas'][j].sum(0)) for j in range(4)]
    ref = oracle.total_area_to_metrics(*tot, ['mIoU', 'mDice', 'mFscore'])
    got = L.IoUMetric.total_area_to_metrics(*tot, ['mIoU', 'mDice', 'mFscore'])
    for k in ref:
        np.testing.assert_array_equal(ref[k], got[k])
        np.testing.assert_array_equal(got[k], g['met_' + k])
    # reference-style 4-tuples are accepted by compute_metrics
    m = L.IoUMetric()
    res = [tuple(torch.from_numpy(g['areas'][j][i]) for j in range(4)) for i in range(3)]
    s = m.compute_metrics(res)
    for k, v in s.items():
        assert float(v) == float(g['sum_' + k])


def test_synth_is_deterministic():
    a = synth.make_images_u8(2, 16, 24, seed=0)
    b = synth.make_images_u8(2, 16, 24, seed=0)
    assert torch.equal(a, b) and a.dtype == torch.uint8 and int(a.max()) <= 254
    lab = synth.make_labels(2, 64, 64, 19, seed=1)
    assert lab.dtype == torch.int64 and set(lab.unique().tolist()) <= set(range(19)) | {255}
    assert 0.02 < (lab == 255).float().mean() < 0.09


def test_engine_param_shapes_cover_engine_expectations():
    from lednet_b200.modules import build_param_shapes
    shapes = build_param_shapes(32, 128, 64, 19)
    assert shapes['decode_head.head_x1.0.conv.weight'] == (19, 32, 3, 3)
    assert shapes['backbone.spp.compression.conv.weight'] == (128, 640, 1, 1)


def test_led_block_modules_mirror_the_reference_surface():
    """SESP / Muti_AFF / GETBBlock / SEAM: registered under the reference's names, state-dict keys equal the oracle's
    (= the reference's, see tests/test_oracle_vs_reference.py), packed parameter blocks have the size the C ABI
    documents, unsupported configurations and CPU tensors fail loudly (no fallback)."""
    from oracle.sesp import OracleSESP
    from oracle.mfaf import OracleMutiAFF
    from oracle.getb import OracleGETBBlock
    from oracle.seam import OracleSEAM
    lib = libmod.get()
    pairs = [
        (L.MODELS.build(dict(type='SESP', nIn=64, nOut=64)), OracleSESP(64, 64)),
        (L.MODELS.build(dict(type='Muti_AFF', channels=64)), OracleMutiAFF(64)),
        (L.MODELS.build(dict(type='GETBBlock', dim=128, num_heads=8, window_size=8)), OracleGETBBlock(128, 8)),
        (L.MODELS.build(dict(type='SEAM', channels=64)), OracleSEAM(64)),
    ]
    for mine, ref in pairs:
        assert set(mine.state_dict()) == set(ref.state_dict()), type(mine).__name__
        assert sum(p.numel() for p in mine.parameters()) == sum(p.numel() for p in ref.parameters())
        mine.load_state_dict(ref.state_dict(), strict=True)
    cpu = torch.device('cpu')
    assert pairs[1][0].packed_params(cpu).numel() == lib.ledb200_mfaf_param_floats(64, 16)
    assert pairs[2][0].packed_params().numel() == lib.ledb200_getb_param_floats(128, 8, 512)
    assert pairs[3][0].packed_params(cpu).numel() == lib.ledb200_seam_param_floats(64)
    assert lib.ledb200_mfaf_workspace_bytes(2, 64) > 0 and lib.ledb200_seam_workspace_bytes(2, 8, 8) >= 2 * 64 * 5
    x = torch.zeros(1, 64, 16, 16)
    for mod, args in ((pairs[0][0], (x,)), (pairs[1][0], (x, x)), (pairs[3][0], (x, x)),
                      (pairs[2][0], (torch.zeros(1, 128, 16, 16),))):
        with pytest.raises(L.LedB200Error):
            mod.eval()(*args)
    for bad in (dict(type='Muti_AFF', channels=60), dict(type='GETBBlock', dim=128, num_heads=8, window_size=4),
                dict(type='SEAM', channels=60), dict(type='SESP', nIn=64, nOut=64, stride=2)):
        with pytest.raises(NotImplementedError):
            L.MODELS.build(bad)
    # C ABI argument checks need no GPU: null buffers / bad shapes are rejected before any launch
    assert lib.ledb200_postprocess(None, 2, 4, 4, None, 0, 4, 4, 0, 0.3, None, 3, None, None) < 0
    assert b'null' in lib.ledb200_last_error()
    assert lib.ledb200_seam_forward(None, None, None, 0, 1, 4, 4, 64, 0.1, None, None, None) < 0
    assert lib.ledb200_getb_forward(None, None, None, 1, 8, 8, None) < 0
