"""Deterministic synthetic weights, images and labels.

numpy's PCG64 streams are stable across platforms and versions, so the same
seed gives the same tensors in the build container and on the GPU box; the same
state dict feeds the CUDA engine, the CPU oracle and (in the build container)
the verbatim reference modules.

* conv / classifier weights: kaiming-normal, fan_out, relu - what
  ``LEDHead.init_weights`` (reference ``decode_heads/led_head.py:53-57``) and mmcv's
  ConvModule apply; classifier biases U(-0.1, 0.1);
* BatchNorm made non-trivial so eval-mode parity means something:
  gamma ~ U(0.5,1.5), beta ~ N(0,0.1), running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5);
* images: uint8 uniform 0..254 like the reference fixture
  (``tests/test_models/test_forward.py:42``);
* labels: int64 in [0,K) from a nearest-upsampled coarse random map ("blocky"),
  with a fraction of pixels set to 255 (ignore).
"""
import zlib

import numpy as np
import torch


def _rng(seed, key):
    return np.random.default_rng([int(seed), zlib.crc32(key.encode())])


def make_state_dict(template, seed=2):
    """template: a state_dict (name -> tensor) giving names and shapes."""
    keys = list(template.keys())
    keyset = set(keys)
    out = {}
    for k in keys:
        t = template[k]
        shape = tuple(t.shape)
        r = _rng(seed, k)
        prefix, _, leaf = k.rpartition('.')
        is_bn = (prefix + '.running_mean') in keyset
        if leaf == 'num_batches_tracked':
            v = np.zeros(shape, dtype=np.int64)
        elif not t.dtype.is_floating_point:           # index buffers (GETB relative_position_index): structural
            v = t.detach().cpu().numpy()
        elif leaf == 'relative_position_bias_table':  # GETB: made non-trivial (reference init is N(0, 0.02))
            v = r.normal(0.0, 0.5, shape)
        elif is_bn:
            if leaf == 'weight':
                v = r.uniform(0.5, 1.5, shape)
            elif leaf == 'bias':
                v = r.normal(0.0, 0.1, shape)
            elif leaf == 'running_mean':
                v = r.normal(0.0, 0.1, shape)
            elif leaf == 'running_var':
                v = r.uniform(0.5, 1.5, shape)
            else:
                raise KeyError(k)
        elif len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            v = r.normal(0.0, np.sqrt(2.0 / fan_out), shape)
        elif len(shape) == 1 and leaf == 'bias':
            v = r.uniform(-0.1, 0.1, shape)
        elif len(shape) == 1:           # PReLU slope
            v = np.full(shape, 0.25)
        else:
            raise KeyError(f'no synthetic rule for {k} {shape}')
        out[k] = torch.from_numpy(np.asarray(v)).to(t.dtype)
    return out


def make_images_u8(n, h, w, seed=0):
    r = _rng(seed, f'img{n}x{h}x{w}')
    return torch.from_numpy(r.integers(0, 255, (n, 3, h, w), dtype=np.uint8))


def make_labels(n, h, w, num_classes, seed=1, ignore_frac=0.05, ignore_index=255, block=32):
    r = _rng(seed, f'lab{n}x{h}x{w}x{num_classes}')
    ch, cw = max(1, -(-h // block)), max(1, -(-w // block))
    coarse = r.integers(0, num_classes, (n, ch, cw), dtype=np.int64)
    lab = np.repeat(np.repeat(coarse, block, axis=1), block, axis=2)[:, :h, :w].copy()
    # per-pixel noise so the confusion matrix is not block-constant
    noise = r.random((n, h, w)) < 0.10
    lab[noise] = r.integers(0, num_classes, int(noise.sum()), dtype=np.int64)
    ign = r.random((n, h, w)) < ignore_frac
    lab[ign] = ignore_index
    return torch.from_numpy(lab)
