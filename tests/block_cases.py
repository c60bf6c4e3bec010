"""Seeded cases shared by tests/golden/make_golden.py (which runs the reference's own MFAF / GETB modules
on them) and the oracle / GPU parity tests: (tag, constructor kwargs, (N, H, W))."""
import torch

from lednet_b200 import synth

MFAF_CASES = [
    ('c64', dict(channels=64), (2, 32, 48)),          # bins nest (H, W multiples of 16)
    ('c64_odd', dict(channels=64), (1, 37, 50)),      # overlapping adaptive-pool bins, nearest up-sampling
    ('c128_small', dict(channels=128), (1, 9, 13)),   # H, W < 16: pooled grids larger than the input
    ('c32_r2', dict(channels=32, r=2), (1, 20, 16)),
]

GETB_CASES = [
    ('d128', dict(dim=128, num_heads=8, window_size=8), (2, 16, 32)),
    ('d128_odd', dict(dim=128, num_heads=8, window_size=8), (1, 13, 21)),     # reflect-padded windows
    ('d256', dict(dim=256, num_heads=8, window_size=8), (1, 16, 24)),
    ('d64_h16', dict(dim=64, num_heads=16, window_size=8, mlp_ratio=2.), (1, 9, 8)),
]

# SEAM edge gate (tools/speed/ddrnet_speed.py:282-338, 388-389): (tag, (N, H, W)); 64 channels (the prototype's conv_1/conv_2)
SEAM_GOLDEN_CHANNELS = [0, 21, 63]
SEAM_CASES = [('s32x48', (2, 32, 48)), ('s37x51', (1, 37, 51)), ('s5x7', (1, 5, 7)), ('s64x96', (1, 64, 96))]


def seam_state_dict(template, seed=23):
    sd = synth.make_state_dict(template, seed=seed)
    sd['fusion_kernel'] = template['fusion_kernel'].detach().clone()      # the constant (0.6, 0.3, 0.1), not a weight
    return sd


def seam_inputs(shape):
    g = torch.Generator().manual_seed(shape[1] * 5 + shape[2])
    x = torch.randn(shape[0], 64, shape[1], shape[2], generator=g)
    xs = torch.randn(shape[0], 64, shape[1], shape[2], generator=g)
    return x, xs


def seam_unstable(e, eps=2e-5, threshold=0.1):
    """[N,1,H,W] bool: pixels whose 0/1 edge mask may legitimately flip under rounding (any of the three Laplacians of
    the normalised edge response `e` within eps of the threshold at the position it is sampled), dilated by conv_2's
    3x3 reach - the mask is a chain of hard thresholds."""
    import torch.nn.functional as F
    k = torch.tensor([-1, -1, -1, -1, 8, -1, -1, -1, -1], dtype=torch.float32).reshape(1, 1, 3, 3)
    lap = lambda s: F.conv2d(e, k, stride=s, padding=1).clamp(min=0)    # noqa: E731
    u = (lap(1) - threshold).abs() < eps
    for s in (2, 4):
        u |= F.interpolate(((lap(s) - threshold).abs() < eps).float(), e.shape[2:], mode='nearest') > 0
    return u, F.max_pool2d(u.float(), 3, 1, 1) > 0


def block_state_dict(template, seed):
    """synth weights; the 1x1 / depthwise convs of these blocks are not followed by ReLU chains, so kaiming
    fan_out scaling keeps activations O(1) (sigmoid / softmax away from saturation)."""
    return synth.make_state_dict(template, seed=seed)


def block_input(case_index, channels, shape, seed0, n_inputs=1):
    g = torch.Generator().manual_seed(seed0 + case_index)
    xs = [torch.randn(shape[0], channels, shape[1], shape[2], generator=g) for _ in range(n_inputs)]
    return xs[0] if n_inputs == 1 else xs


# ---- SegDataPreProcessor / stack_batch cases: (tag, [(h, w)] per sample, size, size_divisor, pad_val, seg_pad_val)
STACK_CASES = [
    ('size', [(37, 52), (40, 64), (33, 64)], (40, 64), None, 0, 255),
    ('divisor', [(37, 52), (45, 50)], None, 32, 0, 255),
    ('divisor1', [(20, 31), (20, 31)], None, 1, 1.5, 7),
    ('size_exact', [(16, 24)], (16, 24), None, 0, 255),
]


def stack_inputs(case_index, shapes, num_classes=19):
    """uint8 BGR images [3,h,w] and int64 label maps [1,h,w] (with some 255 = ignore) per sample."""
    g = torch.Generator().manual_seed(900 + case_index)
    imgs = [torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8) for h, w in shapes]
    labs = []
    for h, w in shapes:
        lab = torch.randint(0, num_classes, (1, h, w), generator=g, dtype=torch.int64)
        lab[torch.rand(1, h, w, generator=g) < 0.05] = 255
        labs.append(lab)
    return imgs, labs


# ---- LED wiring (LEDNet(variant='led') vs the prototype's DDRNet1): (tag, input [n, h, w]); the GETB blocks reflect-pad
#      their input to a multiple of 8 and PyTorch wants that padding smaller than the map, so 1/64 of the input is >= 8
LED_CASES = [('sq512', (1, 512, 512)), ('r512x768', (2, 512, 768))]
LED_GOLDEN_CHANNELS = list(range(0, 128, 16))          # eight of the 128 output channels keep the fixture small


def led_state_dict(template, seed=7):
    sd = synth.make_state_dict(template, seed=seed)
    sd['fusion_kernel'] = template['fusion_kernel'].detach().clone()       # fixed (0.6, 0.3, 0.1) mixing weights, not a weight
    return sd


def led_input(case_index, shape):
    g = torch.Generator().manual_seed(300 + case_index)
    return torch.randn(shape[0], 3, shape[1], shape[2], generator=g)
