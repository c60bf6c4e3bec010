"""Hardware probe (not a test): does a UMMA shared-memory descriptor accept a start address that is
not aligned to the swizzle atom, and an SBO that is not a multiple of the atom?  If so the 3x3 conv
can stage ONE halo slab per Cin chunk (1.4x input fetch) instead of three (3.4x).
Runs each variant in its own process so a trap cannot poison the others."""
import os
import subprocess
import sys

SNIPPET = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0, %r); sys.path.insert(0, %r)
import lednet_b200
from lednet_b200 import ops
g = torch.Generator().manual_seed(1)
for cin, cout, hw in ((64, 64, (32, 40)), (32, 32, (20, 36)), (128, 128, (16, 24))):
    x = torch.randn(2, cin, *hw, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5).bfloat16().float()
    ref = F.conv2d(x, w, None, 1, 1)
    out = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16(), w, None, 1, backend=2)
    err = ((out.float().cpu().permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max()).item()
    print("  %%d->%%d rel err %%.3e %%s" %% (cin, cout, err, "OK" if err < 6e-3 else "WRONG"))
'''
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mode in ('E', 'D0', 'D1'):
    env = dict(os.environ)
    env.pop('LEDB200_TC_LAYOUT', None)
    if mode != 'E':
        env['LEDB200_TC_LAYOUT'] = mode
    print('layout', mode, flush=True)
    try:
        r = subprocess.run([sys.executable, '-c', SNIPPET % (root, os.path.join(root, 'tests'))], env=env,
                           capture_output=True, text=True, timeout=120)
        print(r.stdout[-1500:], r.stderr[-800:] if r.returncode else '', flush=True)
    except subprocess.TimeoutExpired:
        print('  TIMEOUT', flush=True)
