"""-m gpu parity tests of the pretrain_AE step (trainer.py:321-332): weight-gradient GEMM building block, then the
whole step (training forward, fused L1 + backward, clip + Adam) against the oracle's autograd and against the
training fixtures recorded from the LIVE reference (tests/golden/train_*.npz).

Tolerances (fp16 operands and fp16 loss-scaled activation gradients, fp32 accumulation): per-tensor gradient
rel-RMS <= 3e-2 and cosine >= 0.999 against fp32 autograd; loss within 1e-3 relative."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import zs_b200  # noqa: F401
from zs_b200 import _lib, synthetic as syn
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform
from zs_b200 import train as zt
from oracle import ae_oracle as orc
from test_oracle_golden import load_train_golden, train_inputs, sample_idx, TRAIN_CASES
import gpu_helpers as gh

pytestmark = pytest.mark.gpu

GRAD_RELRMS, GRAD_COS = 3e-2, 0.999


# ------------------------------------------------------------------------------------------------
# weight-gradient GEMM (tcgen05, MN-major operands)
# ------------------------------------------------------------------------------------------------
WGRAD_CASES = [
    dict(B=4, T=128, c_out=128, c_in=64, k=1),
    dict(B=3, T=16, c_out=64, c_in=64, k=1),                    # 4 segments per 64-row box, B not a multiple of it
    dict(B=5, T=32, c_out=513, c_in=192, k=1),                  # ragged output channels (decoder linear)
    dict(B=4, T=64, c_out=256, c_in=320, k=3),                  # taps = row offsets, two N tiles
    dict(B=4, T=128, c_out=128, c_in=513, k=4, left=2),         # even kernel of the conv bank (pad k/2, k/2-1)
    dict(B=6, T=64, c_out=128, c_in=128, k=5, stride=2),        # stride 2 through the (parity, pair) view
    dict(B=2, T=32, c_out=256, c_in=64, k=3, ps=True),          # pixel-shuffled layer: dy channel m = r*128 + c
    dict(B=32, T=128, c_out=1024, c_in=1024, k=3),              # full-size decoder conv (split K)
]


@pytest.mark.parametrize('case', WGRAD_CASES, ids=lambda c: '-'.join(f'{k}{v}' for k, v in c.items()))
def test_wgrad_gemm(case):
    torch.manual_seed(0)
    B, T, c_out, c_in, k = case['B'], case['T'], case['c_out'], case['c_in'], case['k']
    stride, ps = case.get('stride', 1), case.get('ps', False)
    left = case.get('left', k // 2)
    dev = 'cuda'
    T_in = T * stride
    dy = (torch.randn(B, T, c_out, device=dev) * 0.5).half()
    halo = 3
    x_rows, x_pitch = gh.round_up(T_in + 2 * halo, 2), gh.round_up(c_in, 8)
    x = torch.full((B, x_rows, x_pitch), float('nan'), dtype=torch.float16, device=dev)
    x[:, :, :c_in] = (torch.randn(B, x_rows, c_in, device=dev)).half()
    dy_halo, dy_pitch = 2, gh.round_up(c_out, 8)
    dyb = torch.zeros(B, T + 2 * dy_halo, dy_pitch, dtype=torch.float16, device=dev)
    dyb[:, dy_halo:dy_halo + T, :c_out] = dy
    grad = torch.zeros(c_out, c_in, k, dtype=torch.float32, device=dev)
    d = _lib.WgradDesc()
    d.dy, d.dy_rows, d.dy_pitch, d.dy_channels, d.dy_ch0, d.dy_row0, d.c_out = dyb.data_ptr(), dyb.shape[1], dy_pitch, dy_pitch, 0, dy_halo, c_out
    d.x, d.x_rows, d.x_pitch, d.x_channels, d.x_ch0, d.x_row0, d.c_in, d.stride = x.data_ptr(), x_rows, x_pitch, x_pitch, 0, halo - left, c_in, stride
    d.B, d.T, d.taps = B, T, k
    d.grad, d.c_in_total, d.ci_off, d.k, d.tap0 = grad.data_ptr(), c_in, 0, k, 0
    d.ps_c, d.scale = (c_out // 2 if ps else 0), 0.25
    _lib.check(_lib.lib().zs_wgrad_cl(C.byref(d), gh.stream()))
    torch.cuda.synchronize()
    # reference: grad[co][ci][j] = 0.25 * sum_{b,t} dy[b,t,co] * x[b, halo-left + stride*t + j, ci]
    ref = torch.zeros(c_out, c_in, k, dtype=torch.float64, device=dev)
    for j in range(k):
        xs = x[:, halo - left + j: halo - left + j + stride * T: stride, :c_in].double()
        ref[:, :, j] = 0.25 * torch.einsum('bto,bti->oi', dy.double(), xs)
    if ps:      # dy channel m = r * c + cc  ->  conv output channel 2 cc + r
        m = torch.arange(c_out, device=dev)
        co = 2 * (m % (c_out // 2)) + m // (c_out // 2)
        ref2 = torch.zeros_like(ref)
        ref2[co] = ref
        ref = ref2
    err = (grad.double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, err


# ------------------------------------------------------------------------------------------------
# the whole step against the oracle / the live-reference fixtures
# ------------------------------------------------------------------------------------------------
def build_train_models(m, dev='cuda'):
    enc_sd, dec_sd = syn.encoder_state_dict(m['seed'], c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2],
                                            enc_size=m['enc_size'], enc_mode='one_hot'), \
        syn.decoder_state_dict(m['seed'], c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'])
    enc = Encoder(c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2], ns=m['ns'], dp=m['dp'],
                  enc_size=m['enc_size'], seg_len=m['seg_len'], enc_mode='one_hot')
    dec = Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'], ns=m['ns'], seg_len=m['seg_len'])
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    return enc.to(dev).train(), dec.to(dev).train()


def cos(a, b):
    return (torch.dot(a.flatten(), b.flatten()) / (a.norm() * b.norm() + 1e-30)).item()


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_train_step_matches_reference(name):
    g = load_train_golden(name)
    m = g['meta']
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    enc, dec = build_train_models(m)
    step = zt.PretrainAE(enc, dec, lr=m['lr'], max_grad_norm=m['max_grad_norm'])
    xd, cd = x.cuda(), c.cuda()
    noise = gumbel_from_uniform(u).cuda()
    km = [k.to(torch.uint8).cuda().contiguous() for k in keep] if keep is not None else None
    step.step_count = 1
    loss, ids = step.forward_backward(xd, cd, noise=noise, keep_masks=km)
    torch.cuda.synchronize()
    # same discrete units as the reference (otherwise the two backward passes differentiate different graphs)
    assert np.array_equal(ids.cpu().numpy(), g['ids']), 'unit ids differ from the reference'
    assert abs(loss.item() - float(g['loss'])) <= 1e-3 * float(g['loss'])

    # (1) against the fixtures recorded from the live reference: norms of every tensor + strided samples
    l_o, ge, gd, _, _ = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, m['dp'], m['ns'], m['seg_len'])
    worst = (0.0, '')
    for net, grads_o, ours in (('enc', ge, step.enc.grad_views), ('dec', gd, step.dec.grad_views)):
        for k, go in grads_o.items():
            gg = ours[k].detach().cpu()
            assert torch.isfinite(gg).all(), (net, k)
            ref_n = float(g[f'gn:{net}:{k}'])
            assert abs(gg.norm().item() - ref_n) <= 3e-2 * ref_n + 1e-7, (net, k, gg.norm().item(), ref_n)
            rs = torch.from_numpy(g[f'g:{net}:{k}'])
            got = gg.reshape(-1)[sample_idx(gg.numel())]
            assert (got - rs).norm().item() <= 5e-2 * rs.norm().item() + 1e-7, (net, k)
            # (2) against the oracle's full gradient tensors
            rel = ((gg - go).norm() / (go.norm() + 1e-30)).item()
            if go.norm().item() > 1e-7:
                assert rel <= GRAD_RELRMS and cos(gg, go) >= GRAD_COS, (net, k, rel, cos(gg, go))
            worst = max(worst, (rel, f'{net}:{k}'))
    print(f'{name}: loss {loss.item():.6f} (ref {float(g["loss"]):.6f}); worst per-tensor gradient rel-RMS {worst[0]:.2e} at {worst[1]}')

    # (3) clip + Adam: the parameters after the update against the reference's
    step._optim(step.enc)
    step._optim(step.dec)
    torch.cuda.synchronize()
    n_enc, n_dec = step.grad_norms()
    assert abs(n_enc - float(g['norm_enc'])) <= 3e-2 * float(g['norm_enc'])
    assert abs(n_dec - float(g['norm_dec'])) <= 3e-2 * float(g['norm_dec'])
    for net, mod in (('enc', enc), ('dec', dec)):
        for k, p in mod.named_parameters():
            got = p.detach().cpu().reshape(-1)[sample_idx(p.numel())].numpy()
            ref_g = g[f'g:{net}:{k}']
            # first Adam step = lr * sign(g) wherever |g| >> eps: must agree unless the gradient is ~0 or its sign is
            # within fp16 noise of flipping
            tol = np.where(np.abs(ref_g) > 1e-6, 2e-5, 2.2 * m['lr'])
            frac_ok = (np.abs(got - g[f'p:{net}:{k}']) <= tol).mean()
            assert frac_ok >= 0.97, (net, k, frac_ok)


def test_train_step_full_batch32_vs_oracle():
    """Config 4 shape (B = 32 per rank): loss and gradients against the oracle's autograd, dropout active."""
    torch.set_num_threads(os.cpu_count())
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, dp=0.5)
    B, T = 32, 128
    enc, dec = build_train_models(m)
    step = zt.PretrainAE(enc, dec)
    x, c = syn.spectrogram_batch(B, T, 5), syn.speaker_ids(B, 102, 5)
    u = syn.gumbel_uniform((B, 16, 1024), 5)
    gen = torch.Generator().manual_seed(11)
    keep = [torch.empty(s).bernoulli_(0.5, generator=gen) for s in orc.dropout_mask_shapes(B, T)]
    km = [k.to(torch.uint8).cuda().contiguous() for k in keep]
    step.step_count = 1
    loss, ids = step.forward_backward(x.cuda(), c.cuda(), noise=gumbel_from_uniform(u).cuda(), keep_masks=km)
    torch.cuda.synchronize()
    enc_sd = {k: v.detach().cpu().clone() for k, v in enc.state_dict().items()}
    dec_sd = {k: v.detach().cpu().clone() for k, v in dec.state_dict().items()}
    l_o, ge, gd, _, ids_o = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, 0.5)
    agree = (ids.cpu().long() == ids_o).float().mean().item()
    assert agree >= 0.95
    assert abs(loss.item() - l_o.item()) <= 2e-3 * l_o.item()
    if agree == 1.0:
        for grads_o, ours in ((ge, step.enc.grad_views), (gd, step.dec.grad_views)):
            for k, go in grads_o.items():
                gg = ours[k].detach().cpu()
                rel = ((gg - go).norm() / (go.norm() + 1e-30)).item()
                assert rel <= GRAD_RELRMS and cos(gg, go) >= GRAD_COS, (k, rel)


def test_training_reduces_loss_and_eval_sees_updates():
    """A few real steps: loss goes down, no overflow skips, and the eval path picks up the updated weights."""
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, dp=0.5)
    enc, dec = build_train_models(m)
    step = zt.PretrainAE(enc, dec, lr=1e-3)
    x, c = syn.spectrogram_batch(8, 128, 2).cuda(), syn.speaker_ids(8, 102, 2).cuda()
    losses = [step.step(x, c).item() for _ in range(12)]
    torch.cuda.synchronize()
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0] - 0.01, losses
    assert step.n_skipped == 0
    enc.eval(); dec.eval()
    noise = gumbel_from_uniform(syn.gumbel_uniform((8, 16, 1024), 0)).cuda()
    act, logits, ids = enc.encode(x, noise)
    spec = dec.decode(act, c)
    with torch.no_grad():
        sd_e = {k: v.cpu() for k, v in enc.state_dict().items()}
        sd_d = {k: v.cpu() for k, v in dec.state_dict().items()}
        a_o, l_o, _ = orc.encoder_forward(sd_e, x.cpu(), syn.gumbel_uniform((8, 16, 1024), 0))
        s_o = orc.decoder_forward(sd_d, act.cpu(), c.cpu())
    assert ((logits.cpu() - l_o).norm() / l_o.norm()).item() < 2e-2
    assert ((spec.cpu() - s_o).norm() / s_o.norm()).item() < 1e-2


def test_autograd_wrappers_match_fused_step():
    """encode_step / decode_step + loss.backward() (the reference's loop shape) give the fused step's gradients."""
    g = load_train_golden('train_small_dp0')
    m = g['meta']
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    enc, dec = build_train_models(m)
    xd, cd, noise = x.cuda(), c.cuda(), gumbel_from_uniform(u).cuda()
    act, _ = zt.encode_step(enc, xd, noise)
    x_dec = zt.decode_step(dec, act, cd)
    loss = torch.mean(torch.abs(x_dec - xd))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g['loss'])) <= 1e-3 * float(g['loss'])
    for net, mod in (('enc', enc), ('dec', dec)):
        for k, p in mod.named_parameters():
            ref_n = float(g[f'gn:{net}:{k}'])
            assert p.grad is not None and abs(p.grad.norm().item() - ref_n) <= 3e-2 * ref_n + 1e-7, (net, k)
