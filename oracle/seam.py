"""Oracle SEAM edge gate (TEST INFRASTRUCTURE) - restates the inline edge path of the authors' speed prototype,
``tools/speed/ddrnet_speed.py``: parameters ``:88-93,102-113``, edge map ``:282-338``, gate ``:388-389``
(``normalize_tensor`` ``:24-37``).  SURVEY section 8(f) rank 1.

    e    = normalize(conv_1(x))                    3x3 conv C -> 1 + BN, min-max normalised over the WHOLE tensor
    b_s  = clamp(laplacian_stride_s(e), 0)         s = 1, 2, 4 ; the strided maps nearest-upsampled to full size
    m    = [0.6 [b_1 > t] + 0.3 [b_2 > t] + 0.1 [b_4 > t]  >  t]        t = boundary_threshold = 0.1
    out  = conv_2(m) * x_s + x_s                   3x3 conv 1 -> C + BN

PINNED: the prototype as a whole cannot be executed here (it needs mmcv / thop and moves a tensor to CUDA in
``__init__``), but its SEAM statements can: ``oracle/ref_loader.load_seam`` slices them out of the file's AST and runs
them verbatim behind the mmcv ConvModule shim.  ``tests/golden/seam.npz`` holds their outputs (make_golden.py),
``tests/test_oracle_golden.py::test_seam_gate`` checks this restatement against them on every run and
``tests/test_oracle_vs_reference.py`` re-checks bit-equality live when ``/root/reference`` is mounted.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .mmcv_shim import ConvModule


class OracleSEAM(nn.Module):
    def __init__(self, channels=64, boundary_threshold=0.1):
        super().__init__()
        norm = dict(type='BN', requires_grad=True)
        self.conv_1 = ConvModule(channels, 1, kernel_size=3, norm_cfg=norm, act_cfg=None, padding=1)      # :102-107
        self.conv_2 = ConvModule(1, channels, kernel_size=3, norm_cfg=norm, act_cfg=None, padding=1)      # :108-113
        self.register_buffer('laplacian_kernel', torch.tensor([-1, -1, -1, -1, 8, -1, -1, -1, -1],
                                                              dtype=torch.float32).reshape(1, 1, 3, 3), persistent=False)
        self.fusion_kernel = nn.Parameter(torch.tensor([[6. / 10], [3. / 10], [1. / 10]],
                                                       dtype=torch.float32).reshape(1, 3, 1, 1), requires_grad=False)
        self.boundary_threshold = boundary_threshold

    def edge_mask(self, x):
        t = self.boundary_threshold
        e = self.conv_1(x)
        e = (e - torch.min(e)) / (torch.max(e) - torch.min(e))                          # normalize_tensor :24-37
        b1 = F.conv2d(e, self.laplacian_kernel, padding=1).clamp(min=0)                 # :286-288
        b1 = (b1 > t).float()                                                           # :292-293
        b2 = F.conv2d(e, self.laplacian_kernel, stride=2, padding=1).clamp(min=0)       # :296-298
        b4 = F.conv2d(e, self.laplacian_kernel, stride=4, padding=1).clamp(min=0)       # :300-302
        b4 = (F.interpolate(b4, b1.shape[2:], mode='nearest') > t).float()              # :304-305, :321-324
        b2 = (F.interpolate(b2, b1.shape[2:], mode='nearest') > t).float()              # :306-307, :313-316
        pyr = F.conv2d(torch.stack((b1, b2, b4), dim=1).squeeze(2), self.fusion_kernel)  # :326-331
        return (pyr > t).float(), e                                                     # :333-338

    def forward(self, x, x_s):
        m, _ = self.edge_mask(x)
        result = self.conv_2(m) * x_s                                                   # :388
        return result + x_s                                                             # :389
