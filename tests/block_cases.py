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


def block_state_dict(template, seed):
    """synth weights; the 1x1 / depthwise convs of these blocks are not followed by ReLU chains, so kaiming
    fan_out scaling keeps activations O(1) (sigmoid / softmax away from saturation)."""
    return synth.make_state_dict(template, seed=seed)


def block_input(case_index, channels, shape, seed0, n_inputs=1):
    g = torch.Generator().manual_seed(seed0 + case_index)
    xs = [torch.randn(shape[0], channels, shape[1], shape[2], generator=g) for _ in range(n_inputs)]
    return xs[0] if n_inputs == 1 else xs
