"""Oracle GETB block (TEST INFRASTRUCTURE) - restates ``GETBBlock`` / ``GlobalLocalAttention`` / ``Mlp`` of
``mmseg/models/backbones/UNetFormer_GETB.py:79-226``, eval mode
(drop / drop_path = 0).

    y = x + attn(BN1(x));  out = y + fc2(ReLU6(fc1(BN2(y))))
    attn(z): reflect-pad z to a multiple of ws -> qkv 1x1 -> per (window, head): softmax(q k^T * scale +
             relative-position bias) v -> crop -> avgpool(ws x 1) + avgpool(1 x ws) over reflect-padded
             rows / columns -> + z -> reflect pad (0,1,0,1) -> depthwise ws x ws conv -> BN -> 1x1 conv -> crop

The einops rearranges of the reference (:174-176, :192-193) are written as view/permute; parameter and buffer
names equal the reference's so one state dict feeds the verbatim module, this oracle and the CUDA kernels.
Third-party: timm (``DropPath``, ``trunc_normal_`` - identity in eval / init only) and einops
(``rearrange`` - pure indexing) are not arithmetic on this path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def relative_position_index(ws):                          # UNetFormer_GETB.py:131-141
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing='ij'))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


class _Conv(nn.Sequential):                               # :35-40
    def __init__(self, cin, cout, k, bias=False):
        super().__init__(nn.Conv2d(cin, cout, k, bias=bias, padding=(k - 1) // 2))


class _SeparableConvBN(nn.Sequential):                    # :56-65
    def __init__(self, cin, cout, k):
        super().__init__(nn.Conv2d(cin, cin, k, padding=(k - 1) // 2, groups=cin, bias=False),
                         nn.BatchNorm2d(cout), nn.Conv2d(cin, cout, 1, bias=False))


class OracleGlobalLocalAttention(nn.Module):              # :97-206
    def __init__(self, dim=256, num_heads=16, qkv_bias=False, window_size=8):
        super().__init__()
        self.num_heads, self.ws = num_heads, window_size
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = _Conv(dim, 3 * dim, 1, bias=qkv_bias)
        self.proj = _SeparableConvBN(dim, dim, window_size)
        ws = window_size
        self.attn_x = nn.AvgPool2d(kernel_size=(ws, 1), stride=1, padding=(ws // 2 - 1, 0))
        self.attn_y = nn.AvgPool2d(kernel_size=(1, ws), stride=1, padding=(0, ws // 2 - 1))
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        self.register_buffer('relative_position_index', relative_position_index(ws))

    def forward(self, x):
        B, C, H, W = x.shape
        ws, h = self.ws, self.num_heads
        local = x
        if W % ws != 0:                                   # :147-154
            x = F.pad(x, (0, ws - W % ws, 0, 0), mode='reflect')
        if H % ws != 0:
            x = F.pad(x, (0, 0, 0, ws - H % ws), mode='reflect')
        Hp, Wp = x.shape[2:]
        hh, ww, d = Hp // ws, Wp // ws, C // h
        qkv = self.qkv(x).view(B, 3, h, d, hh, ws, ww, ws)
        # 'b (qkv h d) (hh ws1) (ww ws2) -> qkv (b hh ww) h (ws1 ws2) d'
        qkv = qkv.permute(1, 0, 4, 6, 2, 5, 7, 3).reshape(3, B * hh * ww, h, ws * ws, d)
        q, k, v = qkv[0], qkv[1], qkv[2]
        dots = (q @ k.transpose(-2, -1)) * self.scale
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(ws * ws, ws * ws, -1)
        dots = dots + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
        attn = dots.softmax(dim=-1) @ v
        # '(b hh ww) h (ws1 ws2) d -> b (h d) (hh ws1) (ww ws2)'
        attn = attn.view(B, hh, ww, h, ws, ws, d).permute(0, 3, 6, 1, 4, 2, 5).reshape(B, C, Hp, Wp)
        attn = attn[:, :, :H, :W]
        out = self.attn_x(F.pad(attn, pad=(0, 0, 0, 1), mode='reflect')) + \
            self.attn_y(F.pad(attn, pad=(0, 1, 0, 0), mode='reflect'))
        out = out + local
        out = F.pad(out, pad=(0, 1, 0, 1), mode='reflect')
        out = self.proj(out)
        return out[:, :, :H, :W]


class _Mlp(nn.Module):                                    # :79-94
    def __init__(self, cin, hidden):
        super().__init__()
        self.fc1 = nn.Conv2d(cin, hidden, 1, 1, 0, bias=True)
        self.act = nn.ReLU6()
        self.fc2 = nn.Conv2d(hidden, cin, 1, 1, 0, bias=True)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class OracleGETBBlock(nn.Module):                         # :209-226
    def __init__(self, dim=256, num_heads=16, mlp_ratio=4., qkv_bias=False, window_size=8):
        super().__init__()
        self.norm1 = nn.BatchNorm2d(dim)
        self.attn = OracleGlobalLocalAttention(dim, num_heads, qkv_bias, window_size)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = nn.BatchNorm2d(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))
