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


def run_cuda_backward(m, x, c, u, keep, d_spec=None):
    """Training forward + backward on the GPU.  d_spec None: fused L1 loss; else the given upstream gradient."""
    enc, dec = build_train_models(m)
    step = zt.PretrainAE(enc, dec, lr=m.get('lr', 1e-4), max_grad_norm=m.get('max_grad_norm', 5.0))
    xd, cd, noise = x.cuda(), c.cuda(), gumbel_from_uniform(u).cuda()
    km = [k.to(torch.uint8).cuda().contiguous() for k in keep] if keep is not None else None
    step.step_count = 1
    if d_spec is None:
        loss, ids = step.forward_backward(xd, cd, noise=noise, keep_masks=km)
    else:
        S = 2.0 ** 15 * x.shape[0]
        act, _, ids = enc.forward_train(xd, noise, 0, km)
        spec = dec.forward_train(act, cd)
        d_act = dec.backward(step.dec.grad_views, S, d_spec=d_spec.cuda())
        enc.backward(d_act, step.enc.grad_views, S, d_act_scale=S)
        loss = (spec - xd).abs().mean()
    torch.cuda.synchronize()
    return step, enc, dec, loss, ids


@pytest.mark.parametrize('name', ['train_small_dp5', 'train_full_b2_dp5'])
def test_backward_weak_kink_config(name):
    """Leaky-relu slope 0.5: exercises the negative-slope branch of every leaky-relu backward while a sign flip between
    the fp16 forward and the fp32 oracle only changes a local gradient by 2x (not 100x), so agreement stays tight."""
    g = load_train_golden(name)
    m = dict(g['meta'], ns=0.5)
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    l_o, ge, gd, spec_o, ids_o = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, m['dp'], 0.5, m['seg_len'])
    step, enc, dec, loss, ids = run_cuda_backward(m, x, c, u, keep, torch.sign(spec_o - x) / x.numel())
    assert torch.equal(ids.cpu().long(), ids_o)
    for net, grads_o, ours in (('enc', ge, step.enc.grad_views), ('dec', gd, step.dec.grad_views)):
        total = torch.sqrt(sum(v.norm() ** 2 for v in grads_o.values())).item()
        for k, go in grads_o.items():
            gg = ours[k].detach().cpu()
            if go.norm().item() < 1e-3 * total:
                continue
            rel = ((gg - go).norm() / go.norm()).item()
            assert rel <= 0.25 and cos(gg, go) >= 0.97, (net, k, rel, cos(gg, go))   # a wrong branch would be O(1) off


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_backward_machinery_smooth_config(name):
    """The STRICT check of every backward kernel.  With leaky-relu slope 1 and the upstream gradient taken from the
    oracle's own sign(x_dec - x) pattern the loss surface has no kinks between the fp16 forward and the fp32 oracle,
    so all 85 gradient tensors must agree tightly (InstanceNorm, pixel shuffle, stride-2, residuals, speaker embeddings,
    GRU BPTT, straight-through softmax, dropout masks replayed from the reference's draws)."""
    g = load_train_golden(name)
    m = dict(g['meta'], ns=1.0)
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    l_o, ge, gd, spec_o, ids_o = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, m['dp'], 1.0, m['seg_len'])
    d_spec = torch.sign(spec_o - x) / x.numel()
    step, enc, dec, loss, ids = run_cuda_backward(m, x, c, u, keep, d_spec)
    assert torch.equal(ids.cpu().long(), ids_o)
    assert abs(loss.item() - l_o.item()) <= 1e-3 * l_o.item()
    worst = (0.0, '')
    for net, grads_o, ours in (('enc', ge, step.enc.grad_views), ('dec', gd, step.dec.grad_views)):
        total = torch.sqrt(sum(v.norm() ** 2 for v in grads_o.values())).item()
        for k, go in grads_o.items():
            gg = ours[k].detach().cpu()
            assert torch.isfinite(gg).all(), (net, k)
            if go.norm().item() < 1e-4 * total:
                # analytically zero (a bias in front of an InstanceNorm): ours must be noise-level too
                assert gg.norm().item() < 1e-3 * total, (net, k, gg.norm().item())
                continue
            rel = ((gg - go).norm() / go.norm()).item()
            assert rel <= GRAD_RELRMS and cos(gg, go) >= GRAD_COS, (net, k, rel, cos(gg, go))
            worst = max(worst, (rel, f'{net}:{k}'))
    print(f'{name} (smooth): worst per-tensor gradient rel-RMS {worst[0]:.2e} at {worst[1]}')


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_train_step_matches_reference(name):
    """The real configuration (leaky-relu slope 0.01, fused L1) against the fixtures recorded from the LIVE reference.

    |x_dec - x| and leaky-relu have kinks: an element whose sign differs between two evaluations of the forward
    changes its local gradient by 2x / 100x, so the per-tensor gradient is ill-conditioned with respect to ANY
    tf32/fp16-class forward - the fp32 oracle itself moves by 3 % (last layer) to 36 % (first layer) per tensor when
    only its weights are rounded to fp16.  That intrinsic floor is measured here and the CUDA gradients have to stay
    within 3x of it; loss, unit ids and gradient norms are held to absolute tolerances."""
    g = load_train_golden(name)
    m = g['meta']
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    step, enc, dec, loss, ids = run_cuda_backward(m, x, c, u, keep)
    # same discrete units as the reference (otherwise the two backward passes differentiate different graphs)
    agree = float((ids.cpu().numpy() == g['ids']).mean())
    assert agree >= 0.95
    assert abs(loss.item() - float(g['loss'])) <= 2e-3 * float(g['loss'])
    if agree < 1.0:
        assert m['emb_size'] < 512, 'unit ids differ from the reference on a full-size fixture'
        pytest.skip(f'{(1 - agree) * 100:.1f} % of the toy fixture\'s unit ids flipped: the two backward passes differentiate '
                    'different graphs (forward agreement and loss were checked)')

    l_o, ge, gd, _, _ = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, m['dp'], m['ns'], m['seg_len'])
    rnd = lambda sd: {k: v.half().float() for k, v in sd.items()}
    _, ge_r, gd_r, _, ids_r = orc.ae_loss_and_grads(rnd(enc_sd), rnd(dec_sd), x, c, u, keep, m['dp'], m['ns'], m['seg_len'])
    report = []
    for net, grads_o, grads_r, ours in (('enc', ge, ge_r, step.enc.grad_views), ('dec', gd, gd_r, step.dec.grad_views)):
        for k, go in grads_o.items():
            gg = ours[k].detach().cpu()
            assert torch.isfinite(gg).all(), (net, k)
            ref_n = float(g[f'gn:{net}:{k}'])            # norm recorded from the live reference
            assert abs(go.norm().item() - ref_n) <= 2e-3 * ref_n + 1e-9
            floor = ((grads_r[k] - go).norm() / go.norm()).item()
            rel = ((gg - go).norm() / go.norm()).item()
            assert rel <= 3.0 * floor + 0.05, (net, k, rel, floor)
            if m['emb_size'] >= 512:       # the 16-channel toy layers of the small fixture are all noise floor
                assert cos(gg, go) >= 0.75, (net, k, cos(gg, go))
            assert abs(gg.norm().item() / ref_n - 1.0) <= 3.0 * floor + 0.1, (net, k, gg.norm().item(), ref_n, floor)
            report.append((rel, floor, f'{net}:{k}'))
    w = max(report)
    print(f'{name}: loss {loss.item():.6f} (ref {float(g["loss"]):.6f}); worst per-tensor gradient rel-RMS {w[0]:.2f} at {w[2]} '
          f'(intrinsic fp16-weight-rounding floor of the oracle there: {w[1]:.2f})')

    # clip + Adam: gradient norms and the parameters after the update against the reference's
    step._optim()
    torch.cuda.synchronize()
    n_enc, n_dec = step.grad_norms()
    assert abs(n_enc - float(g['norm_enc'])) <= 0.3 * float(g['norm_enc'])     # kink noise, see the docstring
    assert abs(n_dec - float(g['norm_dec'])) <= 0.3 * float(g['norm_dec'])
    for net, mod in (('enc', enc), ('dec', dec)):
        for k, p in mod.named_parameters():
            got = p.detach().cpu().reshape(-1)[sample_idx(p.numel())].numpy()
            # first Adam step = -lr * g / (|g| + eps): every weight moves by at most lr, in the reference's direction
            # wherever the gradient sign is stable
            diff = np.abs(got - g[f'p:{net}:{k}'])
            assert (diff <= 2.2 * m['lr']).all(), (net, k)
            stable = np.abs(g[f'g:{net}:{k}']) > 1e-6       # a bias in front of an InstanceNorm has gradient ~0: pure noise
            if stable.sum() >= 16 and m['emb_size'] >= 512:   # (the exact Adam check is test_clip_adam_kernels)
                assert (diff[stable] <= 2e-5).mean() >= 0.6, (net, k, (diff[stable] <= 2e-5).mean())


def test_clip_adam_kernels():
    """zs_grad_sqnorm + zs_adam_step against the oracle's clip_grad_norm / adam_step (utils.py:53-55, trainer.py:64-66),
    three steps, clipping active, plus the overflow skip."""
    torch.manual_seed(0)
    n = 1_000_003
    p0 = torch.randn(n) * 0.05
    params = {'w': p0.clone()}
    state = {}
    pd, md, vd = p0.clone().cuda(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    sq, skipped = torch.zeros(1, device='cuda'), torch.zeros(1, dtype=torch.int32, device='cuda')
    lib = _lib.lib()
    for t in range(1, 4):
        g = torch.randn(n) * (0.02 if t != 2 else 1e-4)           # step 1 and 3 clip (norm 20 > 5), step 2 does not
        norm, gc = orc.clip_grad_norm({'w': g}, 5.0)
        orc.adam_step(params, gc, state, lr=1e-3)
        gd = (g * 2).cuda()                                        # pre-multiplied, undone by grad_mult = 0.5
        sq.zero_()
        _lib.check(lib.zs_grad_sqnorm(gh.ptr(gd), n, gh.ptr(sq), gh.stream()))
        _lib.check(lib.zs_adam_step(gh.ptr(pd), gh.ptr(gd), gh.ptr(md), gh.ptr(vd), n, gh.ptr(sq), 0.5, 5.0, 1e-3, 0.5, 0.9,
                                    1e-8, t, None, gh.ptr(skipped), gh.stream()))
        torch.cuda.synchronize()
        assert abs(sq.sqrt().item() * 0.5 - norm.item()) <= 1e-4 * norm.item()
        assert (pd.cpu() - params['w']).abs().max().item() <= 2e-6
    assert int(skipped.item()) == 0
    before = pd.clone()
    gd[7] = float('inf')
    sq.zero_()
    _lib.check(lib.zs_grad_sqnorm(gh.ptr(gd), n, gh.ptr(sq), gh.stream()))
    _lib.check(lib.zs_adam_step(gh.ptr(pd), gh.ptr(gd), gh.ptr(md), gh.ptr(vd), n, gh.ptr(sq), 0.5, 5.0, 1e-3, 0.5, 0.9, 1e-8, 4,
                                None, gh.ptr(skipped), gh.stream()))
    torch.cuda.synchronize()
    assert int(skipped.item()) == 1 and torch.equal(pd, before)


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
    if agree == 1.0:      # same graph: gradients agree up to the kink noise (see test_train_step_matches_reference)
        for grads_o, ours in ((ge, step.enc.grad_views), (gd, step.dec.grad_views)):
            flat_o = torch.cat([v.flatten() for v in grads_o.values()])
            flat_g = torch.cat([ours[k].detach().cpu().flatten() for k in grads_o])
            assert cos(flat_g, flat_o) >= 0.85
            assert 0.8 <= (flat_g.norm() / flat_o.norm()).item() <= 1.25


def test_side_stream_weight_gradients_equal_in_order_ones():
    """zs_wgrad_async (include/zs_ae.h): weight / bias gradients launched on the library's side streams are the ones the
    in-order launches produce (same kernels, same inputs; only split-K atomic order may differ), eagerly and in the CUDA
    graph, and zs_wgrad_async(0) refuses while gradients are in flight."""
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, dp=0.5)
    B, T = 8, 128
    x, c = syn.spectrogram_batch(B, T, 5).cuda(), syn.speaker_ids(B, 102, 5).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 5)).cuda()
    flats = {}
    for mode in (False, True):
        enc, dec = build_train_models(m)
        step = zt.PretrainAE(enc, dec, async_wgrad=mode, use_graph=False)
        loss, _ = step.forward_backward(x, c, noise=noise, dropout_seed=7)
        torch.cuda.synchronize()
        flats[mode] = (loss.item(), step.enc.grad.clone(), step.dec.grad.clone())
    assert abs(flats[True][0] - flats[False][0]) <= 1e-5 * flats[False][0], (flats[True][0], flats[False][0])   # (atomic partial sums)
    for a, b in zip(flats[True][1:], flats[False][1:]):
        assert torch.isfinite(a).all()
        assert (a - b).norm().item() <= 1e-5 * b.norm().item(), ((a - b).norm().item(), b.norm().item())
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item()
    # graph replays with the fork / join captured.  lr = 0 keeps the parameters fixed, so the gradients of the last replay
    # are comparable (with real updates Adam's early +-lr steps amplify the atomic-order noise of near-zero gradients)
    grads = {}
    for mode in (False, True):
        torch.manual_seed(3)                   # same device-drawn Gumbel noise in both runs
        enc, dec = build_train_models(m)
        step = zt.PretrainAE(enc, dec, lr=0.0, async_wgrad=mode, use_graph=True)
        losses = [step.step(x, c).item() for _ in range(4)]            # two eager steps, capture, two replays
        torch.cuda.synchronize()
        assert step._graph is not None and step.applied_steps() == 4
        grads[mode] = (losses, step.enc.grad.clone(), step.dec.grad.clone())
    for a, b in zip(grads[True][0], grads[False][0]):
        assert abs(a - b) <= 1e-5 * b
    for a, b in zip(grads[True][1:], grads[False][1:]):
        assert b.norm().item() > 0
        assert (a - b).norm().item() <= 1e-5 * b.norm().item(), ((a - b).norm().item(), b.norm().item())
    lib = _lib.lib()
    _lib.check(lib.zs_wgrad_async(1))
    _lib.check(lib.zs_wgrad_join(None))        # nothing pending: no-op
    _lib.check(lib.zs_wgrad_async(0))


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


def test_train_mode_forward_draws_new_dropout_masks_every_call():
    """the drop-in loop (Encoder(x) in train() mode) must not reuse one dropout mask forever.  Two consecutive
    train-mode forwards with the same input and the same Gumbel noise differ (dp > 0); with dp = 0 they are identical; and
    torch.manual_seed pins the masks."""
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, dp=0.5)
    enc, _ = build_train_models(m)
    x = syn.spectrogram_batch(4, 128, 5).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((4, 16, 1024), 5)).cuda()
    torch.manual_seed(7)
    l1 = enc(x, noise)[1].clone()
    l2 = enc(x, noise)[1].clone()
    torch.manual_seed(7)
    l3 = enc(x, noise)[1].clone()
    assert not torch.equal(l1, l2), 'two train-mode forwards used identical dropout masks'
    assert torch.equal(l1, l3), 'torch.manual_seed must pin the dropout masks'
    enc.dp = 0.0
    assert torch.equal(enc(x, noise)[1], enc(x, noise)[1])


def test_pretrain_ae_ranks_and_steps_draw_different_masks():
    """The device-resident seed changes every iteration and with the rank salt (zs_train_meta_begin)."""
    lib = _lib.lib()
    seeds = []
    for salt in (1, 2):
        meta = torch.zeros(8, dtype=torch.int32, device='cuda')
        for _ in range(3):
            _lib.check(lib.zs_train_meta_begin(gh.ptr(meta), salt, None))
            seeds.append(int(meta[0:2].view(torch.int64).item()))
        assert int(meta[6:8].view(torch.int64).item()) == 3
    assert len(set(seeds)) == 6, seeds


def test_overflow_skips_both_networks_and_does_not_count():
    """zs_train_meta_commit: a non-finite norm of EITHER network leaves both untouched and the applied-step count alone."""
    lib = _lib.lib()
    n = 4096
    meta = torch.zeros(8, dtype=torch.int32, device='cuda')
    skipped = torch.zeros(1, dtype=torch.int32, device='cuda')
    p = torch.randn(n, device='cuda'); p0 = p.clone()
    g = torch.randn(n, device='cuda'); mm = torch.zeros(n, device='cuda'); vv = torch.zeros(n, device='cuda')
    sq_ok = (g * g).sum().reshape(1)
    sq_bad = torch.full((1,), float('inf'), device='cuda')
    bc = C.c_void_p(meta.data_ptr() + 8)
    for sq_b, applied in ((sq_bad, 0), (sq_ok, 1), (sq_ok, 2)):
        _lib.check(lib.zs_train_meta_commit(gh.ptr(meta), gh.ptr(sq_ok), gh.ptr(sq_b), 0.5, 0.9, gh.ptr(skipped), None))
        _lib.check(lib.zs_adam_step(gh.ptr(p), gh.ptr(g), gh.ptr(mm), gh.ptr(vv), n, gh.ptr(sq_ok), 1.0, 5.0, 1e-3, 0.5, 0.9, 1e-8,
                                    1, bc, gh.ptr(skipped), None))
        torch.cuda.synchronize()
        assert int(meta[5].item()) == applied
        if applied == 0:
            assert torch.equal(p, p0) and int(skipped.item()) == 1
        else:
            assert not torch.equal(p, p0)
            want_bc1 = 1 - 0.5 ** applied
            assert abs(meta[2:3].view(torch.float32).item() - want_bc1) < 1e-6


def test_autograd_wrappers_refuse_a_stale_backward():
    """the module keeps ONE set of saved activations; a backward of an older forward must raise, not
    silently differentiate the newer forward."""
    g = load_train_golden('train_small_dp0')
    m = g['meta']
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    enc, dec = build_train_models(m)
    xd, cd, noise = x.cuda(), c.cuda(), gumbel_from_uniform(u).cuda()
    act, _ = zt.encode_step(enc, xd, noise)
    y1 = zt.decode_step(dec, act, cd)
    y2 = zt.decode_step(dec, act.detach(), cd)        # second forward of the same module before the first backward
    with pytest.raises(RuntimeError, match='another training forward'):
        y1.abs().mean().backward()
    y2.abs().mean().backward()                        # the latest forward is fine


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
    step, *_ = run_cuda_backward(m, x, c, u, keep)
    for net, mod, fused in (('enc', enc, step.enc.grad_views), ('dec', dec, step.dec.grad_views)):
        for k, p in mod.named_parameters():
            assert p.grad is not None
            ref = fused[k]
            assert (p.grad - ref).norm().item() <= 5e-2 * ref.norm().item() + 1e-6, (net, k)   # different loss scales -> different fp16 roundings


def test_reference_training_loop_runs_unchanged_on_the_modules():
    """trainer.py:321-332 verbatim in shape: Encoder(x) / Decoder(enc_act, c) in train() mode carry an autograd graph,
    `loss.backward()` fills `.grad`, clip_grad_norm_ per network, ONE torch Adam over both networks; the modules
    notice the in-place parameter update and re-pack their tensor-core operands.  Checked against the oracle's step."""
    g = load_train_golden('train_small_dp0')
    m = g['meta']
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    enc, dec = build_train_models(m)
    enc.dp = 0.0
    params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.Adam(params, lr=1e-4, betas=(0.5, 0.9))             # trainer.py:64-66
    xd, cd = x.cuda().requires_grad_(True), c.cuda()                      # utils.py:43-45: to_var -> requires_grad
    enc_act, enc_out = enc(xd, gumbel_from_uniform(u).cuda())             # :325 encode_step
    x_dec = dec(enc_act, cd)                                              # :326 decode_step
    loss_rec = torch.mean(torch.abs(x_dec - xd))                          # :327
    for net in (enc, dec):                                                # :328 reset_grad
        net.zero_grad()
    loss_rec.backward()                                                   # :329
    for net in (enc, dec):                                                # :330 grad_clip per network
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5)
    opt.step()                                                            # :332
    torch.cuda.synchronize()
    assert abs(loss_rec.item() - float(g['loss'])) <= 1e-3 * float(g['loss'])
    o_enc = {k: v.clone() for k, v in enc_sd.items()}
    o_dec = {k: v.clone() for k, v in dec_sd.items()}
    orc.pretrain_ae_step(o_enc, o_dec, {}, x, c, u, keep_masks=None, dp=0.0, ns=m['ns'], seg_len=m['seg_len'])
    moved = agree = 0
    for sd0, sd1, mod in ((enc_sd, o_enc, enc), (dec_sd, o_dec, dec)):
        for k, p in mod.named_parameters():
            want = (sd1[k] - sd0[k])                                      # the oracle's update of this tensor
            got = p.detach().cpu() - sd0[k]
            big = want.abs() > 0.5e-4                                     # first Adam step: |update| ~ lr where |g| >> eps
            moved += int(big.sum())
            agree += int((torch.sign(got[big]) == torch.sign(want[big])).sum())
    # (kinks of |.| and leaky-relu make individual signs ill-conditioned on the toy fixture - see test_train_step_matches_reference)
    assert moved > 1000 and agree >= 0.6 * moved, (agree, moved)
    for mod in (enc, dec):
        for k, p in mod.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
    # the eval path sees the updated weights: the oracle on the module's own (updated) parameters
    enc.eval(); dec.eval()
    with torch.no_grad():
        _, logits2, _ = enc.encode(x.cuda(), gumbel_from_uniform(u).cuda())
        ours = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
        rels = {}
        for name, sd in (('updated', ours), ('original', enc_sd), ('oracle-updated', o_enc)):
            l_o = orc.encoder_forward(sd, x, u, ns=m['ns'], seg_len=m['seg_len'], enc_size=m['enc_size'])[1]
            rels[name] = ((logits2.cpu() - l_o).norm() / l_o.norm()).item()
    print('eval logits rel-RMS vs oracle on', rels)
    assert rels['updated'] < 3e-2, rels


@pytest.mark.parametrize('g_mode', ['naive', 'targeted', 'targeted_residual'])
def test_trainer_steps_gen_step_matches_oracle(g_mode):
    """Trainer.permute_data / encode_step / decode_step / gen_step (trainer.py:238-254, 272-284) in train mode: forward
    value and the Generator's / Decoder's gradients of a loss on x_gen against the oracle's autograd.  Leaky-relu slope 0.5 and
    a squared-error loss keep the comparison away from the sign kinks that make per-element gradients ill-conditioned on the toy
    fixture (see test_train_step_matches_reference); the wiring - which module gets which speaker ids, the combine rule,
    the gradient paths through x_dec and x_dec * G - is what is under test."""
    g = load_train_golden('train_small_dp0')
    m = dict(g['meta'], ns=0.5)
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    enc, dec = build_train_models(m)
    n_spk, n_tgt = m['n_spk'], 2
    mask = g_mode == 'targeted_residual'
    c_a = n_spk if g_mode == 'naive' else n_tgt
    gen_sd = syn.decoder_state_dict(9, c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=c_a)
    gen = Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=c_a, ns=m['ns'], seg_len=m['seg_len'], output_mask=mask)
    gen.load_state_dict(gen_sd)
    gen.cuda().train()
    steps = zt.TrainerSteps(enc, dec, gen, g_mode, n_speakers=n_spk, n_target_speakers=n_tgt)
    c_t = (c % n_tgt) + (n_spk - n_tgt)                       # target-speaker ids, as the stage-2 loaders provide
    C_, X = steps.permute_data((c_t, x.permute(0, 2, 1).contiguous()))
    assert X.shape == x.shape and X.requires_grad and torch.equal(C_.cpu(), c_t)
    enc.dp = 0.0
    enc_act, _ = steps.encode_step(X, gumbel_from_uniform(u).cuda())
    x_gen = steps.gen_step(enc_act.detach(), C_)
    tgt = X.detach()
    loss = torch.mean((x_gen - tgt) ** 2)
    loss.backward()
    torch.cuda.synchronize()
    # oracle: same units (checked), autograd over the restatement
    act_o = enc_act.detach().cpu()
    dec_p = {k: v.clone().requires_grad_(True) for k, v in dec_sd.items()}
    gen_p = {k: v.clone().requires_grad_(True) for k, v in gen_sd.items()}
    x_dec_o = orc.decoder_forward(dec_p, act_o, c_t, m['ns'], m['seg_len'])
    shift = n_spk - n_tgt
    if g_mode == 'naive':
        o = x_dec_o + orc.decoder_forward(gen_p, act_o, c_t, m['ns'], m['seg_len'])
    elif g_mode == 'targeted':
        o = x_dec_o + orc.decoder_forward(gen_p, act_o, c_t - shift, m['ns'], m['seg_len'])
    else:
        o = x_dec_o + x_dec_o * orc.decoder_forward(gen_p, act_o, c_t - shift, m['ns'], m['seg_len'], output_mask=True)
    with torch.no_grad():
        chk = orc.gen_step(dec_sd, gen_sd, act_o, c_t, g_mode, shift, m['ns'], m['seg_len'])
    assert torch.allclose(chk, o.detach(), atol=1e-6)
    assert ((x_gen.detach().cpu() - o.detach()).norm() / o.detach().norm()).item() < 1e-2
    loss_o = torch.mean((o - x) ** 2)
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) <= 2e-3 * loss_o.item()
    tot = lambda gs: torch.sqrt(sum(v.norm() ** 2 for v in gs)).item()
    for mod, ps in ((gen, gen_p), (dec, dec_p)):
        go = torch.cat([ps[k].grad.reshape(-1) for k, _ in mod.named_parameters()])
        gg = torch.cat([p.grad.detach().cpu().reshape(-1) for _, p in mod.named_parameters()])
        assert torch.isfinite(gg).all()
        # whole-network gradient direction and size (per-tensor kink noise: see test_train_step_matches_reference)
        assert cos(gg, go) >= 0.97 and 0.85 <= gg.norm().item() / go.norm().item() <= 1.18, (cos(gg, go), gg.norm().item(), go.norm().item())


def test_twenty_step_loss_curve_follows_the_oracle():
    """20 consecutive pretrain_AE iterations (trainer.py:321-332) at full width against the oracle's step, same
    batches, same Gumbel noise, dropout off: the loss after every update stays within 1 % of the oracle's - i.e. forward, backward,
    per-network clipping, Adam's bias corrections and the in-place operand re-pack all track the fp32 reference over a trajectory,
    not only for one step."""
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, dp=0.0)
    enc, dec = build_train_models(m)
    lr = 1e-3
    step = zt.PretrainAE(enc, dec, lr=lr)
    o_enc = {k: v.detach().cpu().clone() for k, v in enc.state_dict().items()}
    o_dec = {k: v.detach().cpu().clone() for k, v in dec.state_dict().items()}
    state = {}
    torch.set_num_threads(os.cpu_count())
    B, T = 4, 128
    ours, ref = [], []
    for k in range(20):
        x, c = syn.spectrogram_batch(B, T, 300 + k % 3), syn.speaker_ids(B, 102, 300 + k % 3)
        u = syn.gumbel_uniform((B, 16, 1024), 400 + k)
        loss = step.step(x.cuda(), c.cuda(), noise=gumbel_from_uniform(u).cuda(), dropout_seed=0)
        ours.append(loss.item())
        l_o, _, _ = orc.pretrain_ae_step(o_enc, o_dec, state, x, c, u, keep_masks=None, dp=0.0, ns=m['ns'], seg_len=m['seg_len'], lr=lr)
        ref.append(float(l_o))
    print('loss curve (ours / oracle):', ' '.join(f'{a:.4f}/{b:.4f}' for a, b in zip(ours, ref)))
    assert step.n_skipped == 0 and step.applied_steps() == 20
    assert ref[-1] < ref[0] - 0.01                              # the trajectory actually learns
    for k, (a, b) in enumerate(zip(ours, ref)):
        assert abs(a - b) <= 0.01 * b, (k, a, b)
