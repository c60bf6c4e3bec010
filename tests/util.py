import warnings

import torch

import oracle
import lednet_b200 as L
from lednet_b200 import synth


def build_pair(num_classes, seed=2, dtype='fp32', channels=32, ppm=128, head_ch=64):
    """(oracle segmentor on CPU, product EncoderDecoder) sharing one synthetic state dict."""
    torch.manual_seed(0)
    o = oracle.OracleSegmentor(num_classes=num_classes, channels=channels, ppm_channels=ppm,
                               head_channels=head_ch).eval()
    sd = synth.make_state_dict(o.state_dict(), seed=seed)
    o.load_state_dict(sd)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(
            dict(type='LEDNet', channels=channels, ppm_channels=ppm),
            dict(type='LEDHead', in_channels=4 * channels, channels=head_ch, num_classes=num_classes,
                 dropout_ratio=0., tap_channels=channels),
            data_preprocessor=dict(type='SegDataPreProcessor', mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], bgr_to_rgb=True),
            compute_dtype=dtype).eval()
    m.load_state_dict(sd, strict=True)
    return o, m


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def near_tie_mask(logits, tol):
    """pixels whose top-2 logits differ by less than tol * max|logit| (argmax may legitimately flip)."""
    top2 = logits.topk(2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) < tol * logits.abs().max()


def bf16_storage_agreement(o, x, ref_pred):
    """Argmax agreement, against the fp32 oracle, of THE ORACLE ITSELF run with bf16 storage: every
    conv input and every ConvModule / residual-block output rounded to bf16, arithmetic in fp32 -
    what the reference's modules compute under bf16 activations.  On random-init weights the
    top-2 logit margins are not separated from the bf16 rounding noise the way a trained network's
    are, so this (about 99.5 %) is the ceiling any bf16-activation implementation can reach; the
    product path is gated against it."""
    from oracle.mmcv_shim import ConvModule

    def rnd(t):
        return t.bfloat16().float()

    hooks = []
    for mod in o.modules():
        if isinstance(mod, ConvModule) or type(mod).__name__ in ('BasicBlock', 'Bottleneck'):
            hooks.append(mod.register_forward_hook(
                lambda m_, i_, out: rnd(out) if torch.is_tensor(out) else out))
        if isinstance(mod, torch.nn.Conv2d):
            hooks.append(mod.register_forward_pre_hook(lambda m_, i_: (rnd(i_[0]),)))
    try:
        _, pred = o.predict(rnd(x))
    finally:
        for h in hooks:
            h.remove()
    return (pred == ref_pred).float().mean().item()
