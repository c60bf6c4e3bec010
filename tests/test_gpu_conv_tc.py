"""tcgen05/TMEM implicit-GEMM convolution (csrc/conv_tc.cu) against a plain PyTorch fp32 convolution
of the same bf16-rounded operands.  Shapes cover every (kernel, stride, Cin, Cout) class of the
R0 trunk + LEDHead, ragged tiles, resident and streamed weights, 1 and 2 N-tiles."""
import pytest
import torch
import torch.nn.functional as F

from lednet_b200 import ops
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'

CASES = [
    # cin, cout, k, stride, (H, W)
    (32, 32, 3, 1, (20, 36)),       # layer1: 64 B swizzle rows, resident weights
    (64, 64, 3, 1, (32, 40)),       # spatial branch
    (128, 128, 3, 1, (16, 24)),     # context branch, streamed weights
    (256, 256, 3, 1, (16, 8)),      # N = 256
    (32, 19, 3, 1, (24, 40)),       # head_x1: Cout padded to 32
    (128, 64, 3, 1, (16, 32)),      # head
    (128, 64, 1, 1, (16, 24)),      # compression_1
    (640, 128, 1, 1, (16, 32)),     # DAPPM compression (10 chunks)
    (256, 512, 1, 1, (16, 16)),     # two N tiles
    (64, 19, 1, 1, (12, 20)),       # conv_seg
    (32, 32, 3, 2, (32, 48)),       # stem.1
    (64, 128, 3, 2, (32, 32)),      # down_1
    (128, 256, 3, 2, (16, 32)),     # ctx.1.0.conv1 / down_2.1
    (64, 128, 1, 2, (32, 48)),      # downsample 1x1 s2
    (32, 2, 3, 1, (16, 16)),        # K=2 head (Cout padded to 16)
]


@pytest.mark.parametrize('cin,cout,k,stride,hw', CASES)
def test_conv_tc(cin, cout, k, stride, hw):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k)
    n = 2
    x = torch.randn(n, cin, *hw, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x, w, b, stride, k // 2)
    res = torch.randn(ref.shape, generator=g).bfloat16().float()
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    out = ops.conv2d(xd, w, b, stride, relu=False, backend=2)
    assert rel_err(out.float().cpu().permute(0, 3, 1, 2), ref) < 6e-3
    out = ops.conv2d(xd, w, b, stride, relu=True, residual=rd, backend=2)
    assert rel_err(out.float().cpu().permute(0, 3, 1, 2), F.relu(ref + res)) < 6e-3
    # the CUDA-core path on the same operands agrees too (two independent implementations)
    out_d = ops.conv2d(xd, w, b, stride, relu=True, residual=rd, backend=1)
    assert rel_err(out.float(), out_d.float()) < 8e-3


@pytest.mark.parametrize('cin,cout,stride,nhw', [
    (64, 64, 1, (4, 128, 256)),      # 7 tiles per CTA, one slab per tile: two MMA issuers alternate tiles
    (32, 32, 1, (2, 256, 256)),      # 64 B swizzle rows, alternate-tile epilogue groups
    (128, 64, 1, (4, 128, 128)),     # two Cin chunks per tile, 147 KB of resident weights (head conv)
    (64, 128, 2, (8, 128, 256)),     # stride 2: six slabs per tile on a shorter ring -> single issuer
    (128, 128, 1, (4, 64, 128)),     # streamed weights
])
def test_conv_tc_many_tiles_persistent(cin, cout, stride, nhw):
    # more tiles than SMs: exercises the persistent loop, the TMEM accumulator stages, ring wrap-around and the
    # hand-over between the two MMA issuer threads (with and without a residual)
    g = torch.Generator().manual_seed(7 + cin + cout)
    x = torch.randn(nhw[0], cin, nhw[1], nhw[2], generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).bfloat16().float()
    ref = F.conv2d(x, w, None, stride, 1)
    res = torch.randn(ref.shape, generator=g).bfloat16().float()
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    out = ops.conv2d(xd, w, None, stride, backend=2)
    assert rel_err(out.float().cpu().permute(0, 3, 1, 2), ref) < 6e-3
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    out = ops.conv2d(xd, w, None, stride, relu=True, residual=rd, backend=2)
    assert rel_err(out.float().cpu().permute(0, 3, 1, 2), F.relu(ref + res)) < 6e-3


@pytest.mark.parametrize('hw', [(64, 96), (70, 94), (34, 50)])
@pytest.mark.parametrize('raw_u8', [False, True])
def test_stem_tc_matches_oracle(hw, raw_u8):
    """csrc/stem_tc.cu (thread-built im2col + tcgen05) against the oracle's stem[0] ConvModule and the
    head_x1 pre-activation BN+ReLU of it (second output), for float and raw-uint8 input."""
    import oracle
    from lednet_b200 import synth
    from util import build_pair
    o, m = build_pair(19, dtype='bf16')
    img = synth.make_images_u8(2, *hw, seed=3)
    x = oracle.preprocess(img)
    with torch.no_grad():
        ref_x1 = o.backbone.stem[0](x)
        ref_x1h = o.decode_head.head_x1[0].activate(o.decode_head.head_x1[0].bn(ref_x1))
    eng = m.engine()
    eng.forward_infer(img.to(DEV) if raw_u8 else x.to(DEV))
    assert eng.op_info()[0][1] == 'conv_tc'            # the stem ran on the tensor-core kernel
    got_x1, got_x1h = eng.debug_fetch('x1'), eng.debug_fetch('x1h')
    assert rel_err(got_x1, ref_x1) < 1.5e-2
    assert rel_err(got_x1h, ref_x1h) < 1.5e-2
