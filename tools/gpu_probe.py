"""Bring-up probe for the GPU box: runs the building blocks and the full path, printing error
magnitudes stage by stage (a diagnostic, not a test).  Usage: python tools/gpu_probe.py [stage ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import zs_b200  # noqa: E402
from zs_b200 import _lib, synthetic as syn  # noqa: E402
from zs_b200.model import Encoder, Decoder, gumbel_from_uniform  # noqa: E402
import gpu_helpers as gh  # noqa: E402
from conftest import load_golden  # noqa: E402
from oracle import ae_oracle as orc  # noqa: E402


def stage_conv():
    torch.manual_seed(0)
    cases = [
        dict(B=2, C_in=64, C_out=128, T=32, k=1),
        dict(B=2, C_in=64, C_out=128, T=128, k=1),
        dict(B=3, C_in=128, C_out=256, T=64, k=3),
        dict(B=4, C_in=512, C_out=512, T=128, k=5),
        dict(B=4, C_in=512, C_out=512, T=128, k=5, stride=2),
        dict(B=5, C_in=513, C_out=130, T=77, k=3, lrelu=True, inorm=True),
        dict(B=33, C_in=96, C_out=513, T=16, k=1, lrelu=True),
        dict(B=2, C_in=1024, C_out=1024, T=207, k=3, lrelu=True, inorm=True),
        dict(B=3, C_in=64, C_out=128, T=51, k=5, stride=2, inorm=True),
    ]
    for cs in cases:
        B, C_in, C_out, T, k = cs['B'], cs['C_in'], cs['C_out'], cs['T'], cs['k']
        kw = {a: cs[a] for a in ('stride', 'lrelu', 'inorm') if a in cs}
        x = torch.randn(B, C_in, T, device='cuda')
        W = torch.randn(C_out, C_in, k, device='cuda') / (C_in * k) ** 0.5
        b = torch.randn(C_out, device='cuda') * 0.1
        y = gh.conv_cl(x, W, b, **kw)
        ref = gh.conv_ref(x, W, b, **kw)
        err = (y - ref).abs().max().item()
        print(f'conv {cs}: max|err|={err:.3e} ref_rms={ref.pow(2).mean().sqrt().item():.3f} nan={torch.isnan(y).sum().item()}',
              flush=True)


def _weights(m):
    enc_sd = syn.encoder_state_dict(m['seed'], c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2],
                                    enc_size=m['enc_size'], enc_mode=m['enc_mode'])
    dec_sd = syn.decoder_state_dict(m['seed'], c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'])
    return enc_sd, dec_sd


def stage_model(names):
    for name in names:
        g = load_golden(name)
        m = g['meta']
        enc_sd, dec_sd = _weights(m)
        enc = Encoder(c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2], ns=m['ns'], dp=0.5,
                      enc_size=m['enc_size'], seg_len=m['seg_len'], enc_mode=m['enc_mode'])
        dec = Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'], ns=m['ns'],
                      seg_len=m['seg_len'])
        enc.load_state_dict(enc_sd, strict=True)
        dec.load_state_dict(dec_sd, strict=True)
        enc.cuda().eval()
        dec.cuda().eval()
        x = syn.spectrogram_batch(m['B'], m['T'], m['seed'], c_in=m['c_in']).cuda()
        c = syn.speaker_ids(m['B'], m['n_spk'], m['seed']).cuda()
        noise = gumbel_from_uniform(torch.from_numpy(g['uniform'])) if 'uniform' in g else None
        t0 = time.time()
        act, logits, ids = enc.encode(x, noise)
        torch.cuda.synchronize()
        lg = torch.from_numpy(g['logits'])
        le = (logits.cpu() - lg).abs().max().item()
        lr = ((logits.cpu() - lg).pow(2).mean().sqrt() / lg.pow(2).mean().sqrt()).item()
        agree = (act.cpu().argmax(1).numpy() == g['act_argmax']).mean()
        # decode from the REFERENCE's activations so decoder error is measured on identical units
        if m['enc_mode'] == 'one_hot':
            ref_act = torch.zeros_like(act.cpu()).scatter_(1, torch.from_numpy(g['act_argmax']).long().unsqueeze(1), 1.0)
        else:
            ref_act = torch.from_numpy(g['act'].astype(np.float32))
        spec = dec(ref_act.cuda(), c)
        torch.cuda.synchronize()
        sg = torch.from_numpy(g['spec'])
        se = (spec.cpu() - sg).abs().max().item()
        sr = ((spec.cpu() - sg).pow(2).mean().sqrt() / sg.pow(2).mean().sqrt()).item()
        print(f'{name}: logits max|err|={le:.3e} relrms={lr:.3e} unit-agree={agree * 100:.2f}% | '
              f'spec max|err|={se:.3e} relrms={sr:.3e} nan={torch.isnan(spec).sum().item()} ({time.time() - t0:.2f}s)',
              flush=True)


def main():
    stages = sys.argv[1:] or ['check', 'conv', 'small', 'full']
    print('device:', torch.cuda.get_device_name(0), flush=True)
    if 'check' in stages:
        print('zs_device_check:', _lib.lib().zs_device_check(), _lib.lib().zs_last_error(), flush=True)
    if 'conv' in stages:
        stage_conv()
    if 'small' in stages:
        stage_model(['small_onehot', 'small_onehot_odd', 'small_mbv', 'small_continues', 'small_gumbel_t'])
    if 'full' in stages:
        stage_model(['full_b2_t128', 'full_b1_t207', 'full_b1_t9', 'full_b1_mbv', 'full_b1_e512'])


if __name__ == '__main__':
    main()
