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


def class_palette(num_classes):
    """One BGR colour per class on the 3x3x3 grid {40,127,214}^3 (27 colours; classes beyond wrap with a shade offset)."""
    lv = np.array([40, 127, 214], dtype=np.int64)
    pal = np.zeros((num_classes, 3), dtype=np.int64)
    for k in range(num_classes):
        j = (k * 7 + 3) % 27          # stride 7 walks the whole grid (gcd(7,27)=1): neighbouring classes differ a lot
        pal[k] = np.array([lv[j % 3], lv[(j // 3) % 3], lv[j // 9]]) + 12 * (k // 27)
    return pal


def make_scene(n, h, w, num_classes, seed=0, coarse=(32, 32), noise=24, ignore_frac=0.0, ignore_index=255):
    """A learnable synthetic segmentation scene: a coarse random class map nearest-upsampled to (h, w)
    ("blocky" labels) and an image whose pixels are the class colour plus uniform noise.
    A network trained on such scenes for a few hundred steps has TRAINED logit margins (confident inside
    regions, small only along region borders), which is what north_star's >= 99.9 % bf16 argmax-agreement
    gate presumes; random-init weights do not (tests/util.py).  Returns (uint8 BGR [n,3,h,w], int64 [n,h,w])."""
    r = _rng(seed, f'scene{n}x{h}x{w}x{num_classes}')
    ch, cw = coarse
    cmap = r.integers(0, num_classes, (n, ch, cw), dtype=np.int64)
    ys = (np.arange(h) * ch // h)[:, None]
    xs = (np.arange(w) * cw // w)[None, :]
    lab = cmap[:, ys, xs]                                             # [n,h,w]
    pal = class_palette(num_classes)
    img = pal[lab]                                                    # [n,h,w,3]
    img = img + r.integers(-noise, noise + 1, img.shape, dtype=np.int64)
    img = np.clip(img, 0, 255).astype(np.uint8).transpose(0, 3, 1, 2).copy()
    lab = lab.copy()
    if ignore_frac > 0:
        lab[r.random((n, h, w)) < ignore_frac] = ignore_index
    return torch.from_numpy(img), torch.from_numpy(lab)
