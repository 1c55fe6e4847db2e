import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import zs_b200
from zs_b200 import _lib
import gpu_helpers as gh
case, mode, fn = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
cs = [dict(B=5, C_in=513, C_out=130, T=77, k=3), dict(B=40, C_in=64, C_out=256, T=256, k=3), dict(B=33, C_in=96, C_out=513, T=16, k=1)][case]
torch.manual_seed(1)
xx = torch.randn(cs['B'], cs['C_in'], cs['T'], device='cuda')
W = torch.randn(cs['C_out'], cs['C_in'], cs['k'], device='cuda') / (cs['C_in'] * cs['k']) ** 0.5
bb = torch.randn(cs['C_out'], device='cuda') * 0.1
_lib.lib().zs_set_epilogue_mode(mode)
if fn == 'cl':
    y = gh.conv_cl_to_cl(xx, W, bb, lrelu=True, inorm=True, halo_out=2)
else:
    y = gh.conv_cl(xx, W, bb, lrelu=True, act=1)
print('ok', case, mode, fn, float(y.abs().mean()))
