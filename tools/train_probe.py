"""Per-tensor gradient error table of the CUDA pretrain_AE step against the oracle's autograd (debug aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import zs_b200
from zs_b200 import train as zt
from zs_b200.model import gumbel_from_uniform
from oracle import ae_oracle as orc
from test_oracle_golden import load_train_golden, train_inputs
from test_gpu_train import build_train_models

name = sys.argv[1] if len(sys.argv) > 1 else 'train_small_dp0'
g = load_train_golden(name); m = g['meta']
if os.environ.get('PROBE_NS'): m['ns'] = float(os.environ['PROBE_NS'])
torch.set_num_threads(os.cpu_count())
enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
enc, dec = build_train_models(m)
step = zt.PretrainAE(enc, dec)
km = [k.to(torch.uint8).cuda().contiguous() for k in keep] if keep is not None else None
step.step_count = 1
l_o, ge, gd, spec_o, _ = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, m['dp'], m['ns'], m['seg_len'])
if os.environ.get('PROBE_L1'):
    loss, ids = step.forward_backward(x.cuda(), c.cuda(), noise=gumbel_from_uniform(u).cuda(), keep_masks=km)
else:   # same sign pattern as the oracle's L1 gradient: isolates the backward machinery from sign flips of |spec - x|
    S = 2.0 ** 15 * x.shape[0]
    step.enc.grad.zero_(); step.dec.grad.zero_()
    act, _, ids = enc.forward_train(x.cuda(), gumbel_from_uniform(u).cuda(), 0, km)
    spec = dec.forward_train(act, c.cuda())
    d_spec = (torch.sign(spec_o - x) / x.numel()).cuda()
    d_act = dec.backward(step.dec.grad_views, S, d_spec=d_spec)
    enc.backward(d_act, step.enc.grad_views, S, d_act_scale=S)
    loss = (spec - x.cuda()).abs().mean()
torch.cuda.synchronize()
print('ids equal', np.array_equal(ids.cpu().numpy(), g['ids']), 'loss', loss.item(), 'ref', float(g['loss']))
for net, grads_o, ours in (('dec', gd, step.dec.grad_views), ('enc', ge, step.enc.grad_views)):
    for k, go in grads_o.items():
        gg = ours[k].detach().cpu()
        rel = ((gg - go).norm() / (go.norm() + 1e-30)).item()
        cs = (torch.dot(gg.flatten(), go.flatten()) / (gg.norm() * go.norm() + 1e-30)).item()
        flag = '' if rel < 3e-2 else '   <<<<'
        print(f'{net}:{k:28s} |ref| {go.norm().item():.3e} |ours| {gg.norm().item():.3e} rel {rel:.3e} cos {cs:.5f}{flag}')
