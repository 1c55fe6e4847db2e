"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle and the
live-reference golden fixtures.  Tolerances (fp16 operands = tf32-class 10-bit mantissa, fp32 accumulate;
SURVEY.md section 0 fact 6 / BASELINE.md section 4):
  * unit ids: bit-exact given identical logits and noise; end-to-end agreement >= 95 %
  * logits: rel-RMS <= 1e-2 ; decoded spectrogram (same units): rel-RMS <= 1e-2, max-abs <= 3e-2
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import zs_b200  # noqa: F401
from zs_b200 import _lib, synthetic as syn
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform
from zs_b200.frontend import AutoencoderPath, segment_plan
from oracle import ae_oracle as orc
from conftest import load_golden
import gpu_helpers as gh

pytestmark = pytest.mark.gpu

LOGIT_RELRMS, SPEC_RELRMS, SPEC_MAXABS = 1e-2, 1e-2, 3e-2


def relrms(a, b):
    return ((a - b).norm() / b.norm()).item()


def build_models(m, dev='cuda'):
    enc_sd = syn.encoder_state_dict(m['seed'], c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2],
                                    enc_size=m['enc_size'], enc_mode=m['enc_mode'])
    dec_sd = syn.decoder_state_dict(m['seed'], c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'])
    enc = Encoder(c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2], ns=m['ns'], dp=0.5,
                  enc_size=m['enc_size'], seg_len=m['seg_len'], enc_mode=m['enc_mode'])
    dec = Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'], ns=m['ns'],
                  seg_len=m['seg_len'])
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    return enc.to(dev).eval(), dec.to(dev).eval(), enc_sd, dec_sd


FULL = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, enc_mode='one_hot', emb_size=1024, n_spk=102,
            ns=0.01, seg_len=128)


@pytest.fixture(scope='module')
def full_models():
    return build_models(FULL)


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    dict(B=2, C_in=64, C_out=128, T=32, k=1),
    dict(B=3, C_in=128, C_out=256, T=64, k=3),
    dict(B=4, C_in=512, C_out=512, T=128, k=5),
    dict(B=4, C_in=512, C_out=512, T=128, k=5, stride=2),
    dict(B=5, C_in=513, C_out=130, T=77, k=3, lrelu=True, inorm=True),
    dict(B=33, C_in=96, C_out=513, T=16, k=1, lrelu=True),
    dict(B=2, C_in=1024, C_out=1024, T=207, k=3, lrelu=True, inorm=True),
    dict(B=3, C_in=64, C_out=128, T=51, k=5, stride=2, inorm=True),
    dict(B=1, C_in=8, C_out=8, T=9, k=5),
    dict(B=40, C_in=64, C_out=256, T=256, k=3, lrelu=True, inorm=True),
]


@pytest.mark.parametrize('cs', CONV_CASES, ids=lambda c: '-'.join(f'{k}{v}' for k, v in c.items()))
@pytest.mark.parametrize('operand', ['fp16', 'bf16'])
def test_conv_gemm_matches_operand_rounded_reference(cs, operand):
    torch.manual_seed(1)
    kw = {a: cs[a] for a in ('stride', 'lrelu', 'inorm') if a in cs}
    x = torch.randn(cs['B'], cs['C_in'], cs['T'], device='cuda')
    W = torch.randn(cs['C_out'], cs['C_in'], cs['k'], device='cuda') / (cs['C_in'] * cs['k']) ** 0.5
    b = torch.randn(cs['C_out'], device='cuda') * 0.1
    y = gh.conv_cl(x, W, b, operand=operand, **kw)
    ref = gh.conv_ref(x, W, b, operand=operand, **kw)
    assert not torch.isnan(y).any()
    # fp32 accumulation-order noise only (the one-pass InstanceNorm variance adds a little on long segments)
    assert (y - ref).abs().max().item() < (3e-4 if cs.get('inorm') else 1e-4)


def test_conv_gemm_tile_shapes_agree():
    """Different segments-per-tile choices (nb) must give the same numbers (tiles only regroup columns)."""
    torch.manual_seed(2)
    x = torch.randn(9, 128, 32, device='cuda')
    W = torch.randn(256, 128, 3, device='cuda') / 20
    b = torch.zeros(256, device='cuda')
    base = gh.conv_cl(x, W, b, inorm=True, lrelu=True, nb_hint=1)
    for nb in (2, 3, 8):
        assert torch.equal(gh.conv_cl(x, W, b, inorm=True, lrelu=True, nb_hint=nb), base)


@pytest.mark.parametrize('B,T,nb', [(7, 64, 3), (50, 16, 12), (9, 48, 5), (300, 32, 0)])
def test_conv_gemm_rounds_do_not_spill_across_tiles(B, T, nb):
    """Segments-per-tile that the epilogue's 128-frame store rounds do not divide (regression: a partial round's
    TMA store box used to overwrite the next tile's segments)."""
    torch.manual_seed(5)
    x = torch.randn(B, 64, T, device='cuda')
    W = torch.randn(128, 64, 3, device='cuda') / 14
    b = torch.randn(128, device='cuda') * 0.1
    ref = gh.conv_ref(x, W, b, lrelu=True, inorm=True)
    for _ in range(3):
        y = gh.conv_cl_to_cl(x, W, b, lrelu=True, inorm=True, nb_hint=nb)
        assert (y - ref).abs().max().item() < 2e-3       # fp16 output rounding


def test_bottleneck_argmax_bit_exact_with_ties():
    torch.manual_seed(3)
    B, Cn, T8 = 7, 1024, 16
    logits = torch.randn(B, Cn, T8)
    logits[0, 5, 3] = logits[0, 900, 3] = 50.0          # exact tie -> first index
    logits[1, :, 0] = 0.25                               # all equal -> index 0
    u = torch.rand(B, T8, Cn)
    noise = gumbel_from_uniform(u)
    noise[0, 3, :] = 0.0
    noise[1, 0, :] = 0.0
    act = torch.empty(B, Cn, T8, device='cuda')
    ids = torch.empty(B, T8, dtype=torch.int32, device='cuda')
    logits_d, noise_d = logits.cuda(), noise.cuda()       # keep the device copies alive across the launch
    _lib.check(_lib.lib().zs_bottleneck_one_hot(gh.ptr(logits_d), gh.ptr(noise_d), B, Cn, T8, gh.ptr(act),
                                                gh.ptr(ids), gh.stream()))
    torch.cuda.synchronize()
    want = (logits.permute(0, 2, 1) + noise).max(dim=-1)[1]
    assert torch.equal(ids.cpu().long(), want)
    assert ids[0, 3].item() == 5 and ids[1, 0].item() == 0
    assert torch.equal(act.cpu(), torch.zeros(B, T8, Cn).scatter_(-1, want.unsqueeze(-1), 1.0).permute(0, 2, 1))
    # and the reference's softmax-then-max formulation gives the same ids away from exp-rounding merges
    hard, ind = orc.gumbel_hard(logits[2:].permute(0, 2, 1), u[2:])
    assert torch.equal(ind, want[2:])


@pytest.mark.parametrize('H,B,T,impl', [(128, 5, 16, 1), (128, 5, 16, 2), (512, 3, 40, 1), (512, 3, 40, 2),
                                        (512, 37, 128, 2), (64, 20, 9, 2), (32, 9, 7, 0), (256, 16, 33, 2)])
def test_gru_recurrence_matches_oracle(H, B, T, impl):
    torch.manual_seed(4)
    w_hh = (torch.rand(2, 3 * H, H) * 2 - 1) / H ** 0.5
    b_hh = (torch.rand(2, 3 * H) * 2 - 1) / H ** 0.5
    gx = torch.randn(B, T, 2, 3 * H).half().float()     # the projections are an operand-type (fp16) buffer
    out = torch.zeros(B, T, 2 * H, dtype=torch.float16, device='cuda')
    gx_d, w_d, b_d = gx.cuda(), w_hh.cuda(), b_hh.cuda()
    _lib.check(_lib.lib().zs_gru_recurrence(gh.ptr(gx_d), gh.ptr(w_d), gh.ptr(b_d), B, T, H,
                                            gh.ptr(out), T, 2 * H, 0, 0, 0, impl, gh.stream()))
    torch.cuda.synchronize()
    # oracle: feed both directions' projections as 6H input channels and let W_ih select its half
    x = gx.reshape(B, T, 6 * H).permute(0, 2, 1)
    eye, zero = torch.eye(3 * H), torch.zeros(3 * H, 3 * H)
    sd = {'RNN.weight_ih_l0': torch.cat([eye, zero], 1), 'RNN.weight_ih_l0_reverse': torch.cat([zero, eye], 1),
          'RNN.weight_hh_l0': w_hh[0], 'RNN.weight_hh_l0_reverse': w_hh[1],
          'RNN.bias_ih_l0': torch.zeros(3 * H), 'RNN.bias_ih_l0_reverse': torch.zeros(3 * H),
          'RNN.bias_hh_l0': b_hh[0], 'RNN.bias_hh_l0_reverse': b_hh[1]}
    ref = orc.bi_gru(x, sd)                                  # (B, 2H, T)
    got = out.float().cpu().permute(0, 2, 1)
    assert (got - ref).abs().max().item() < 2e-3            # fp16 store of h in (-1, 1)


# ------------------------------------------------------------------------------------------------
# whole path vs golden fixtures of the live reference
# ------------------------------------------------------------------------------------------------
GOLDEN_GPU = ['small_onehot', 'small_onehot_odd', 'small_mbv', 'small_continues', 'small_gumbel_t', 'small_binary', 'small_zeropad',
              'full_b2_t128', 'full_b1_t207', 'full_b1_mbv', 'full_b1_e512']


@pytest.mark.parametrize('name', GOLDEN_GPU)
def test_path_matches_reference_golden(name):
    g = load_golden(name)
    m = g['meta']
    enc, dec, _, _ = build_models(m)
    x = syn.spectrogram_batch(m['B'], m['T'], m['seed'], c_in=m['c_in']).cuda()
    c = syn.speaker_ids(m['B'], m['n_spk'], m['seed']).cuda()
    u = torch.from_numpy(g['uniform']) if 'uniform' in g else None
    noise = gumbel_from_uniform(u) if u is not None else None
    act, logits, ids = enc.encode(x, noise)
    ref_logits = torch.from_numpy(g['logits'])
    assert relrms(logits.cpu(), ref_logits) < (3e-2 if m['c_h'][1] < 512 else LOGIT_RELRMS)
    if m['enc_mode'] == 'one_hot':
        # bit-exact given identical logits + noise: redo the reference's op on OUR logits on the CPU
        _, want = orc.gumbel_hard(logits.cpu().permute(0, 2, 1), u)
        assert torch.equal(ids.cpu().long(), want)
        assert torch.equal(act.cpu().argmax(1), want)
        agree = (ids.cpu().numpy() == g['act_argmax']).mean()
        assert agree >= 0.9, agree       # 16-26 unit frames per fixture: allow one near-tie flip
        ref_act = torch.zeros_like(act.cpu()).scatter_(1, torch.from_numpy(g['act_argmax']).long().unsqueeze(1), 1.0)
    else:
        ref_act = torch.from_numpy(g['act'].astype(np.float32))
        if m['enc_mode'] == 'continues':
            assert relrms(act.cpu(), ref_act) < 3e-2
        else:
            assert (act.cpu() == ref_act).float().mean().item() >= 0.97
    spec = dec(ref_act.cuda(), c)        # decoder error on identical units
    ref_spec = torch.from_numpy(g['spec'])
    assert spec.shape == ref_spec.shape
    assert relrms(spec.cpu(), ref_spec) < SPEC_RELRMS
    assert (spec.cpu() - ref_spec).abs().max().item() < SPEC_MAXABS
    if m['enc_mode'] == 'one_hot':
        # unit-id gather path == dense one-hot GEMM path, bit for bit
        spec_ids = dec.decode(None, c, unit_ids=torch.from_numpy(g['act_argmax']).cuda())
        assert torch.equal(spec_ids, spec)
    if m['patch']:                       # trainer.py:208-209 'targeted' patcher, fused accumulate
        gen_sd = syn.decoder_state_dict(m['seed'] + 7, c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=2)
        gen = Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=2, ns=m['ns'], seg_len=m['seg_len'])
        gen.load_state_dict(gen_sd)
        gen.cuda().eval()
        c_t = torch.from_numpy(g['c_target']).cuda()
        y = dec(ref_act.cuda(), c_t)
        gen.decode(ref_act.cuda(), c_t - (m['n_spk'] - 2), out=y, accumulate=1)
        assert relrms(y.cpu(), torch.from_numpy(g['spec_patched'])) < SPEC_RELRMS


def test_nine_frame_segment(full_models):
    """MIN_LEN segment (convert.py:36): InstanceNorm over 2-3 frames is ill-conditioned, so only the structure and
    the decoder (on the reference's units) are held to tolerance."""
    g = load_golden('full_b1_t9')
    enc, dec, _, _ = full_models
    x = syn.spectrogram_batch(1, 9, 0).cuda()
    noise = gumbel_from_uniform(torch.from_numpy(g['uniform']))
    act, logits, ids = enc.encode(x, noise)
    assert logits.shape == (1, 1024, 2) and not torch.isnan(logits).any()
    ref_act = torch.zeros(1, 1024, 2).scatter_(1, torch.from_numpy(g['act_argmax']).long().unsqueeze(1), 1.0)
    spec = dec(ref_act.cuda(), syn.speaker_ids(1, 102, 0).cuda())
    assert spec.shape == (1, 513, 16)
    assert relrms(spec.cpu(), torch.from_numpy(g['spec'])) < SPEC_RELRMS


# ------------------------------------------------------------------------------------------------
# whole path vs the oracle on seeded inputs, and size-independent properties at full size
# ------------------------------------------------------------------------------------------------
def test_full_size_batch_vs_oracle(full_models):
    enc, dec, enc_sd, dec_sd = full_models
    B, T = 8, 128
    x = syn.spectrogram_batch(B, T, 11)
    c = syn.speaker_ids(B, 102, 11)
    u = syn.gumbel_uniform((B, 16, 1024), 11)
    act, logits, ids = enc.encode(x.cuda(), gumbel_from_uniform(u))
    spec = dec(act, c.cuda())
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        o_act, o_logits, o_ids = orc.encoder_forward(enc_sd, x, u)
        o_spec = orc.decoder_forward(dec_sd, act.cpu(), c)
    assert relrms(logits.cpu(), o_logits) < LOGIT_RELRMS
    assert (ids.cpu().long() == o_ids).float().mean().item() >= 0.95
    assert relrms(spec.cpu(), o_spec) < SPEC_RELRMS
    assert (spec.cpu() - o_spec).abs().max().item() < SPEC_MAXABS


def test_batching_is_exact_and_segments_independent(full_models):
    """B = 32 x 128 frames (BASELINE config 2): every segment's result equals its own batch-1 result, in any
    batch order - the reference's per-chunk loop (convert.py:154-165) and the batched call are interchangeable."""
    enc, dec, _, _ = full_models
    B, T = 32, 128
    x = syn.spectrogram_batch(B, T, 5).cuda()
    c = syn.speaker_ids(B, 102, 5).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 5))
    act, logits, ids = enc.encode(x, noise)
    spec = dec.decode(None, c, unit_ids=ids)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    act_p, logits_p, ids_p = enc.encode(x[perm.cuda()].contiguous(), noise[perm])
    assert torch.equal(logits_p, logits[perm.cuda()])
    assert torch.equal(ids_p, ids[perm.cuda()])
    spec_p = dec.decode(None, c[perm.cuda()], unit_ids=ids_p)
    assert torch.equal(spec_p, spec[perm.cuda()])
    for b in (0, 13, 31):
        a1, l1, i1 = enc.encode(x[b:b + 1].contiguous(), noise[b:b + 1])
        assert torch.equal(l1[0], logits[b]) and torch.equal(i1[0], ids[b])
        assert torch.equal(dec.decode(None, c[b:b + 1], unit_ids=i1)[0], spec[b])
    assert float(spec.min()) > 0.0 and float(spec.max()) < 1.0        # sigmoid output
    assert torch.equal(act.sum(1), torch.ones(B, 16, device='cuda'))    # one unit per frame


@pytest.mark.parametrize('B', [300, 470, 960])
def test_large_batches_match_small_batches(full_models, B):
    """Saturating batches give bit-identical results to 32-segment batches, run after run.  Above 240 segments the decoder
    GRU switches to the wide kernel with W_hh's r|z rows in TMEM: B = 300 runs gru_wide_kernel<., 2> (64 sequences per
    cluster), B = 470 is where the cost rule of launch_gru_cluster picks gru_wide_kernel<., 4> (128 sequences per cluster),
    B = 960 is the batch bench.py times."""
    enc, dec, _, _ = full_models
    T = 128
    x = syn.spectrogram_batch(B, T, 21).cuda()
    c = syn.speaker_ids(B, 102, 21).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 21)).cuda()
    act, logits, ids = enc.encode(x, noise)
    spec = dec.decode(None, c, unit_ids=ids)
    for s0 in range(0, B, 32):
        a1, l1, i1 = enc.encode(x[s0:s0 + 32].contiguous(), noise[s0:s0 + 32])
        assert torch.equal(l1, logits[s0:s0 + 32]) and torch.equal(i1, ids[s0:s0 + 32])
        assert torch.equal(dec.decode(None, c[s0:s0 + 32], unit_ids=i1), spec[s0:s0 + 32])
    act2, logits2, ids2 = enc.encode(x, noise)
    assert torch.equal(logits2, logits)
    assert torch.equal(dec.decode(None, c, unit_ids=ids2), spec)


@pytest.mark.parametrize('H,B,T', [(512, 70, 24), (512, 300, 16), (512, 470, 12), (128, 300, 16)])
@pytest.mark.parametrize('operand', ['fp16', 'bf16'])
def test_gru_kernel_variants_agree_with_oracle(H, B, T, operand):
    """Every tensor-core GRU instantiation the forward passes can select - gru_cluster_kernel (16 / 32 sequences) and
    gru_wide_kernel<., 2 | 4> - in BOTH operand types, against the oracle's explicit recurrence."""
    torch.manual_seed(8)
    dt = torch.float16 if operand == 'fp16' else torch.bfloat16
    tol = 2e-3 if operand == 'fp16' else 1.6e-2          # the state is stored in the operand type: 2^-11 vs 2^-8 relative
    w_hh = ((torch.rand(2, 3 * H, H) * 2 - 1) / H ** 0.5).to(dt).float()
    b_hh = (torch.rand(2, 3 * H) * 2 - 1) / H ** 0.5
    gx = torch.randn(B, T, 2, 3 * H).to(dt).float()
    out = torch.zeros(B, T, 2 * H, dtype=dt, device='cuda')
    gx_d, w_d, b_d = gx.cuda(), w_hh.cuda(), b_hh.cuda()
    _lib.check(_lib.lib().zs_gru_recurrence(gh.ptr(gx_d), gh.ptr(w_d), gh.ptr(b_d), B, T, H, gh.ptr(out), T, 2 * H, 0, 0,
                                            _lib.OPERANDS[operand], 0, gh.stream()))
    torch.cuda.synchronize()
    x = gx.reshape(B, T, 6 * H).permute(0, 2, 1)
    eye, zero = torch.eye(3 * H), torch.zeros(3 * H, 3 * H)
    sd = {'RNN.weight_ih_l0': torch.cat([eye, zero], 1), 'RNN.weight_ih_l0_reverse': torch.cat([zero, eye], 1),
          'RNN.weight_hh_l0': w_hh[0], 'RNN.weight_hh_l0_reverse': w_hh[1],
          'RNN.bias_ih_l0': torch.zeros(3 * H), 'RNN.bias_ih_l0_reverse': torch.zeros(3 * H),
          'RNN.bias_hh_l0': b_hh[0], 'RNN.bias_hh_l0_reverse': b_hh[1]}
    torch.set_num_threads(os.cpu_count())
    ref = orc.bi_gru(x, sd)
    got = out.float().cpu().permute(0, 2, 1)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < tol, (got - ref).abs().max().item()


def test_whole_path_in_bf16_operands():
    """operand = 'bf16' (the range-safe switch for checkpoints whose activations exceed fp16): the whole encode -> decode
    path against the oracle at the precision SURVEY fact 6 measured for bf16 operands (unit agreement ~98 %,
    spectrogram rel-RMS ~2e-2)."""
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, enc_mode='one_hot')
    enc, dec, enc_sd, dec_sd = build_models(m)
    enc.operand = dec.operand = 'bf16'
    B, T = 16, 128
    x, c = syn.spectrogram_batch(B, T, 31), syn.speaker_ids(B, 102, 31)
    u = syn.gumbel_uniform((B, 16, 1024), 31)
    act, logits, ids = enc.encode(x.cuda(), gumbel_from_uniform(u))
    spec = dec(act, c.cuda())
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        _, o_logits, o_ids = orc.encoder_forward(enc_sd, x, u)
        o_spec = orc.decoder_forward(dec_sd, act.cpu(), c)
    agree = (ids.cpu().long() == o_ids).float().mean().item()
    print(f'bf16 operands: unit agreement {agree * 100:.1f} %, logits rel-RMS {relrms(logits.cpu(), o_logits):.2e}, '
          f'spectrogram rel-RMS {relrms(spec.cpu(), o_spec):.2e}')
    assert agree >= 0.93 and relrms(logits.cpu(), o_logits) < 8e-2 and relrms(spec.cpu(), o_spec) < 4e-2


def test_fp16_and_frames_major_inputs_are_bit_identical(full_models):
    """zs_encoder_forward_x: an fp16 upload and the (B, T, 513) layout of Trainer.test_step's argument (trainer.py:196) give
    exactly the results of the fp32 (B, 513, T) call; ids-only calls (no logits / act buffers) give the same ids."""
    enc, dec, _, _ = full_models
    for B, T in ((5, 128), (3, 77), (2, 207)):
        x = syn.spectrogram_batch(B, T, 61).cuda()
        noise = gumbel_from_uniform(syn.gumbel_uniform((B, Encoder.t8(T), 1024), 61)).cuda()
        act, logits, ids = enc.encode(x, noise)
        for xin, layout in ((x.half(), 'nct'), (x.permute(0, 2, 1).contiguous(), 'ntc'), (x.permute(0, 2, 1).contiguous().half(), 'ntc')):
            a2, l2, i2 = enc.encode(xin, noise, layout=layout)
            assert torch.equal(l2, logits) and torch.equal(i2, ids) and torch.equal(a2, act), (T, layout, xin.dtype)
        a3, l3, i3 = enc.encode(x, noise, want_act=False, want_logits=False)
        assert a3 is None and l3 is None and torch.equal(i3, ids)


def test_device_drawn_gumbel_noise(full_models):
    """Throughput mode: noise=None + per-segment seeds -> the bottleneck draws its Gumbel noise in the kernel.  (1) the draw of
    a segment depends on its seed only (not on the batch it rides in); (2) the draws are Gumbel(0, 1): with all-zero logits
    the winner is uniform over the units and repeats at the birthday rate; (3) the host replay of the generator gives the
    same ids through the explicit-noise path (bit-exact argmax given identical logits and noise)."""
    enc, _, _, _ = full_models
    B, T = 6, 128
    x = syn.spectrogram_batch(B, T, 71).cuda()
    seeds = torch.arange(1, B + 1, dtype=torch.int64, device='cuda') * 7919
    _, _, ids = enc.encode(x, None, noise_seeds=seeds, want_act=False, want_logits=False)
    perm = torch.tensor([4, 2, 0, 5, 1, 3], device='cuda')
    _, _, ids_p = enc.encode(x[perm].contiguous(), None, noise_seeds=seeds[perm].contiguous(), want_act=False, want_logits=False)
    assert torch.equal(ids_p, ids[perm])
    _, _, ids_other = enc.encode(x, None, noise_seeds=seeds + 1, want_act=False, want_logits=False)
    assert not torch.equal(ids_other, ids)
    # host replay of the counter-based generator (csrc/kernels.cuh gumbel_from_counter)
    M = (1 << 64) - 1

    def gumbel_host(seed, n):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        with np.errstate(over='ignore'):
            z = np.uint64(seed & M) + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        return u
    us = np.stack([gumbel_host(int(s), 16 * 1024).reshape(16, 1024) for s in seeds.cpu().tolist()])
    u_t = torch.from_numpy(us)
    assert 0.49 < float(u_t.mean()) < 0.51 and float(u_t.min()) >= 0.0 and float(u_t.max()) < 1.0
    _, logits, ids_host = enc.encode(x, gumbel_from_uniform(u_t))
    # logf on the device vs torch.log on the host differ in the last ulp: ids agree except at near-ties
    assert (ids_host == ids).float().mean().item() >= 0.97
    # distribution: zero logits -> argmax of pure Gumbel noise is uniform over 1024 units
    z = torch.zeros(64, 1024, 16, device='cuda')
    ids0 = torch.empty(64, 16, dtype=torch.int32, device='cuda')
    s64 = torch.arange(64, dtype=torch.int64, device='cuda') * 104729 + 5
    lib = _lib.lib()
    # (drive the kernel through the encoder entry point's building block: explicit zero logits need the unit-test entry)
    g = torch.from_numpy(np.stack([gumbel_host(int(s), 16 * 1024).reshape(16, 1024) for s in s64.cpu().tolist()]))
    _lib.check(lib.zs_bottleneck_one_hot(gh.ptr(z), gh.ptr(gumbel_from_uniform(g).cuda()), 64, 1024, 16, None, gh.ptr(ids0), gh.stream()))
    torch.cuda.synchronize()
    counts = torch.bincount(ids0.flatten().long().cpu(), minlength=1024)
    assert counts.max().item() <= 8 and (counts > 0).sum().item() > 600        # 1024 draws over 1024 bins


def test_fp16_saturation_is_counted_and_raised():
    """the fp16 clamp must not be silent.  conv3's weights x 3e5 push its (un-normalised) output past
    65504: the epilogues count it, `check_range` raises OperandRangeError naming the fix, and operand = 'bf16' runs clean."""
    from zs_b200.model import OperandRangeError, check_range
    m = dict(seed=0, c_in=513, c_h=[128, 512, 128], enc_size=1024, emb_size=1024, n_spk=102, ns=0.01, seg_len=128, enc_mode='one_hot')
    enc, dec, enc_sd, dec_sd = build_models(m)
    x = syn.spectrogram_batch(2, 128, 3).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((2, 16, 1024), 3))
    check_range()                                  # reset
    enc.encode(x, noise)
    assert check_range() == 0                      # the synthetic / reference value ranges stay far inside fp16
    with torch.no_grad():
        enc.conv3.weight.mul_(3e5)                 # conv3 feeds conv4 directly: no normalisation in between
    enc.encode(x, noise)
    with pytest.raises(OperandRangeError, match="bf16"):
        check_range()
    assert check_range() == 0                      # the counter was reset by the check
    enc.operand = 'bf16'
    _, logits, _ = enc.encode(x, noise)
    assert check_range() == 0 and torch.isfinite(logits).all()


def test_frontend_matches_per_chunk_reference_loop(full_models):
    """convert()/encode() semantics on ragged utterances: chunk plan, MIN_LEN padding/truncation, tail chunk."""
    enc, dec, enc_sd, dec_sd = full_models
    path = AutoencoderPath(enc, dec, seg_len=128, max_batch=16)
    rng = np.random.Generator(np.random.PCG64(7))
    lengths = [5, 9, 100, 128, 130, 300, 391]
    specs = [np.clip(rng.random((L, 513), dtype=np.float32), 1e-8, 1) for L in lengths]
    spk = [3, 7, 100, 101, 0, 55, 20]
    torch.manual_seed(99)
    outs, units = path.convert_utterances(specs, spk, enc_only=True, reference_noise_order=True)
    # replay the reference's loop chunk by chunk with the same generator state, through the oracle on CPU
    torch.manual_seed(99)
    for u, spec in enumerate(specs):
        padded, plan, keep = segment_plan(len(spec), 128)
        sp = np.concatenate([spec, np.zeros((padded - len(spec), 513), np.float32)]) if padded > len(spec) else spec
        assert outs[u].shape[0] == sum(8 * Encoder.t8(e - s) for s, e in plan)
        n_units = sum(Encoder.t8(e - s) for s, e in plan)
        assert units[u].shape == ((keep if keep is not None else n_units), 1024)
        assert set(np.unique(units[u])) <= {0.0, 1.0}
    # one utterance end-to-end against the oracle (noise order = chunk order)
    torch.manual_seed(99)
    noises = []
    for u, spec in enumerate(specs):
        padded, plan, keep = segment_plan(len(spec), 128)
        noises.append([torch.rand(1, Encoder.t8(e - s), 1024) for s, e in plan])
    u = 5
    _, plan, _ = segment_plan(lengths[u], 128)
    agree = []
    with torch.no_grad():
        for (s, e), un in zip(plan, noises[u]):
            xs = torch.from_numpy(specs[u][s:e]).t().unsqueeze(0)
            _, _, o_ids = orc.encoder_forward(enc_sd, xs, un)
            agree.append(o_ids[0])
    o_ids = torch.cat(agree)
    got = torch.from_numpy(units[u]).argmax(1)
    assert (got == o_ids).float().mean().item() >= 0.9


def test_errors_are_loud(full_models):
    enc, dec, _, _ = full_models
    with pytest.raises(RuntimeError):
        enc(torch.rand(1, 513, 8, device='cuda'))          # below MIN_LEN
    with pytest.raises(RuntimeError):
        enc(torch.rand(1, 513, 300, device='cuda'))        # longer than 2*seg_len
    with pytest.raises(RuntimeError):
        enc(torch.rand(1, 100, 128, device='cuda'))        # wrong channel count
    with pytest.raises(RuntimeError):
        dec(torch.rand(1, 1024, 16, device='cuda'), torch.tensor([102]))   # speaker out of range (host ids)
    with pytest.raises(RuntimeError):                       # zero-padding mode (seg_len < 64) is inference-only
        e = Encoder(ns=0.01, enc_size=32, seg_len=32, enc_mode='one_hot', c_in=33, c_h1=16, c_h2=64, c_h3=16).cuda().train()
        e.forward_train(torch.rand(1, 33, 40, device='cuda'), torch.zeros(1, 5, 32, device='cuda'))


def test_streaming_host_to_host_overlapping_calls(full_models):
    """StreamingResynthesizer: pinned host in -> pinned host out in micro-batches, two calls in flight at once
    (buffer sets rotate across calls); results equal the direct device-resident calls bit for bit."""
    from zs_b200.frontend import StreamingResynthesizer
    enc, dec, _, _ = full_models
    S, T = 50, 128
    st = StreamingResynthesizer(enc, dec, micro_batch=16, n_buffers=3)
    calls = []
    for k in range(3):
        x = syn.spectrogram_batch(S, T, 40 + k).pin_memory()
        c = syn.speaker_ids(S, 102, 40 + k).pin_memory()
        noise = gumbel_from_uniform(syn.gumbel_uniform((S, 16, 1024), 40 + k)).pin_memory()
        spec = torch.zeros(S, 513, T).pin_memory()
        ids = torch.zeros(S, 16, dtype=torch.int32).pin_memory()
        calls.append((x, c, noise, spec, ids, st.run_async(x, c, spec, ids, noise)))     # no host wait in between
    for call in calls:
        call[-1].synchronize()       # a module handle takes ONE call in flight: finish the stream before the direct calls
    for x, c, noise, spec, ids, done in calls:
        _, _, ids_d = enc.encode(x.cuda(), noise.cuda())
        spec_d = dec.decode(None, c.cuda(), unit_ids=ids_d)
        assert torch.equal(ids, ids_d.cpu())
        assert torch.equal(spec, spec_d.cpu())
    # the stream-ordered form: results are visible to work queued on the caller's stream after run()
    x, c, noise, spec, ids, _ = calls[0]
    spec.zero_()
    st.run(x, c, spec, ids, noise)
    torch.cuda.current_stream().synchronize()
    assert float(spec.min()) > 0.0


def test_zero_padding_mode_edges():
    """seg_len < 64 selects 'constant' padding (model/model.py:36-38).  The speaker embedding is folded into per-speaker
    bias tables, which under zero padding needs the edge corrections -W_0.e / -W_2.e at the first / last frame of every
    k = 3 decoder conv: check the segment edges specifically, and that they are where the two padding modes differ."""
    g = load_golden('small_zeropad')
    m = g['meta']
    assert m['seg_len'] < 64
    enc, dec, enc_sd, dec_sd = build_models(m)
    c = syn.speaker_ids(m['B'], m['n_spk'], m['seed'])
    ref_act = torch.zeros(m['B'], m['enc_size'], g['act_argmax'].shape[1]).scatter_(
        1, torch.from_numpy(g['act_argmax']).long().unsqueeze(1), 1.0)
    spec = dec(ref_act.cuda(), c.cuda()).cpu()
    ref = torch.from_numpy(g['spec'])                         # the live reference, zero padding
    with torch.no_grad():
        reflect = orc.decoder_forward(dec_sd, ref_act, c, ns=m['ns'], seg_len=128)
    edge = [0, 1, 2, 3, -4, -3, -2, -1]
    assert (reflect[:, :, edge] - ref[:, :, edge]).abs().max().item() > 5e-2      # the modes really differ at the edges
    assert (spec[:, :, edge] - ref[:, :, edge]).abs().max().item() < SPEC_MAXABS
    assert relrms(spec[:, :, edge], ref[:, :, edge]) < SPEC_RELRMS


@pytest.mark.parametrize('enc_size', [512, 1024])
def test_long_utterance_with_patcher(enc_size):
    """BASELINE config 5: AE + TTS patcher (second Decoder, g_mode 'targeted', trainer.py:200-209) on a 2000-frame
    utterance -> chunks 14 x 128 + 1 x 207 (convert.py:154-165), enc_size 512 and 1024, against the oracle's
    test_step chunk by chunk with the reference's noise order."""
    m = dict(FULL, enc_size=enc_size)
    enc, dec, enc_sd, dec_sd = build_models(m)
    gen_sd = syn.decoder_state_dict(7, c_in=enc_size, c_out=513, c_h=1024, c_a=2)
    gen = Decoder(c_in=enc_size, c_out=513, c_h=1024, c_a=2, ns=0.01, seg_len=128)
    gen.load_state_dict(gen_sd)
    path = AutoencoderPath(enc, dec, gen, g_mode='targeted', n_speakers=102, n_target_speakers=2, seg_len=128, max_batch=16)
    rng = np.random.Generator(np.random.PCG64(11))
    spec = np.clip(rng.random((2000, 513), dtype=np.float32), 1e-8, 1)
    _, plan, _ = segment_plan(2000, 128)
    assert [e - s for s, e in plan] == [128] * 14 + [207]
    torch.manual_seed(5)
    outs, units = path.convert_utterances([spec], [101], enc_only=False, reference_noise_order=True)
    assert outs[0].shape == (14 * 128 + 8 * Encoder.t8(207), 513) and units[0].shape == (14 * 16 + Encoder.t8(207), enc_size)
    torch.manual_seed(5)
    noises = [torch.rand(1, Encoder.t8(e - s), enc_size) for s, e in plan]
    got_units = torch.from_numpy(units[0]).argmax(1)
    row = urow = 0
    agree, n_units = 0, 0
    with torch.no_grad():
        for (s, e), un in list(zip(plan, noises))[::5] + [(plan[-1], noises[-1])]:      # every 5th chunk + the 207-frame tail
            k = plan.index((s, e))
            urow, row = sum(Encoder.t8(b - a) for a, b in plan[:k]), sum(8 * Encoder.t8(b - a) for a, b in plan[:k])
            xs = torch.from_numpy(spec[s:e]).t().unsqueeze(0)
            T8 = Encoder.t8(e - s)
            _, _, o_ids = orc.encoder_forward(enc_sd, xs, un, enc_size=enc_size)
            ours = got_units[urow:urow + T8]
            agree += int((ours == o_ids[0]).sum()); n_units += T8
            # decoder + patcher on OUR units (isolates the decoder/patcher error from unit flips)
            act = torch.zeros(1, enc_size, T8).scatter_(1, ours.view(1, 1, T8), 1.0)
            c = torch.tensor([101])
            want = orc.decoder_forward(dec_sd, act, c) + orc.decoder_forward(gen_sd, act, c - 100)     # trainer.py:209
            got = torch.from_numpy(outs[0][row:row + 8 * T8]).t().unsqueeze(0)
            assert relrms(got, want) < SPEC_RELRMS and (got - want).abs().max().item() < 2 * SPEC_MAXABS
    assert agree >= 0.95 * n_units, (agree, n_units)


@pytest.mark.parametrize('g_mode', ['naive', 'targeted', 'targeted_residual'])
def test_trainer_test_step_surface(full_models, g_mode):
    """`AutoencoderPath.test_step / encoder_test_step` = Trainer.test_step / encoder_test_step (trainer.py:194-228): same
    argument layout ((B, T, 513) in), numpy out, the three Decoder-as-patcher combine rules (trainer.py:206-211) fused into
    the generator's output epilogue, and the target-speaker guard (trainer.py:202-203)."""
    enc, dec, enc_sd, dec_sd = full_models
    residual = g_mode == 'targeted_residual'
    n_gen = 102 if g_mode == 'naive' else 2
    gen_sd = syn.decoder_state_dict(9, c_in=1024, c_out=513, c_h=1024, c_a=n_gen)
    gen = Decoder(c_in=1024, c_out=513, c_h=1024, c_a=n_gen, ns=0.01, seg_len=128, output_mask=residual)
    gen.load_state_dict(gen_sd)
    path = AutoencoderPath(enc, dec, gen, g_mode=g_mode, n_speakers=102, n_target_speakers=2, seg_len=128)
    B, T = 2, 128
    x = syn.spectrogram_batch(B, T, 31).permute(0, 2, 1).contiguous()           # (B, T, 513) as convert_x hands it over
    c = torch.tensor([100, 101])
    u = syn.gumbel_uniform((B, 16, 1024), 31)
    noise = gumbel_from_uniform(u)
    x_dec, enc_np = path.test_step(x, c, enc_only=False, noise=noise)
    assert isinstance(x_dec, np.ndarray) and x_dec.shape == (B, 513, T) and enc_np.shape == (B, 1024, 16)
    enc_only, _ = path.test_step(x, c, enc_only=True, noise=noise)
    assert np.array_equal(path.encoder_test_step(x, noise=noise), enc_np)
    ids = torch.from_numpy(enc_np).argmax(1)
    with torch.no_grad():
        act = torch.zeros(B, 1024, 16).scatter_(1, ids.unsqueeze(1), 1.0)           # our units: isolates the decoders' error
        base = orc.decoder_forward(dec_sd, act, c)
        if g_mode == 'naive':
            want = base + orc.decoder_forward(gen_sd, act, c)
        elif g_mode == 'targeted':
            want = base + orc.decoder_forward(gen_sd, act, c - 100)
        else:
            want = base + base * orc.decoder_forward(gen_sd, act, c - 100, output_mask=True)
        o_ids = orc.encoder_forward(enc_sd, x.permute(0, 2, 1), u)[2]
    assert (ids == o_ids).float().mean().item() >= 0.95
    assert relrms(torch.from_numpy(enc_only), base) < SPEC_RELRMS
    assert relrms(torch.from_numpy(x_dec), want) < SPEC_RELRMS
    assert (torch.from_numpy(x_dec) - want).abs().max().item() < 2 * SPEC_MAXABS
    if g_mode != 'naive':
        with pytest.raises(RuntimeError):
            path.test_step(x, torch.tensor([3, 101]), enc_only=False, noise=noise)   # not a target speaker


# ------------------------------------------------------------------------------------------------
# alternate TTS patchers (model/model.py:492-552) and their Trainer combine rule
# ------------------------------------------------------------------------------------------------
from test_oracle_golden import PATCHER_CASES, patcher_inputs       # noqa: E402
from conftest import load_golden as _load_golden                   # noqa: E402


def build_patcher(m, sd):
    from zs_b200.patchers import Enhanced_Generator, Spectrogram_Patcher
    if m['kind'] == 'spectrogram':
        net = Spectrogram_Patcher(c_in=m['c_in'], c_out=m['c_in'], c_h=m['c_h'], c_a=m['c_a'], ns=0.01, seg_len=128)
    else:
        net = Enhanced_Generator(ns=0.01, dp=0.5, enc_size=m['enc_size'], emb_size=m['emb_size'], seg_len=128, n_speakers=m['n_spk'])
        if m['c_in'] != 513:      # narrow fixture: the reference hard-codes the widths, the fixture script rebuilt the parts
            net.Encoder = Encoder(c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2], ns=0.01, dp=0.5,
                                  enc_size=m['enc_size'], seg_len=128, enc_mode='continues')
            net.Decoder = Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'], ns=0.01, seg_len=128)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


@pytest.mark.parametrize('name', PATCHER_CASES)
def test_patchers_match_reference_golden(name):
    g = _load_golden(name)
    m = g['meta']
    sd, x, c = patcher_inputs(g)
    net = build_patcher(m, sd)
    assert list(net.state_dict().keys()) == list(sd.keys())          # checkpoint contract: same names, same order
    y = net(x.cuda(), c.cuda())
    ref = torch.from_numpy(g['out'])
    assert relrms(y.cpu(), ref) < SPEC_RELRMS and (y.cpu() - ref).abs().max().item() < SPEC_MAXABS
    # fused combine rule: x_dec += Generator(x_dec, c) with the output aliasing the input (trainer.py:212-213)
    xd = x.cuda().clone()
    net.patch(xd, c.cuda(), out=xd, accumulate=1)
    assert torch.allclose(xd.cpu(), x + y.cpu(), atol=1e-6)


@pytest.mark.parametrize('g_mode', ['enhanced', 'spectrogram'])
def test_trainer_test_step_with_alternate_patchers(full_models, g_mode):
    """Trainer.test_step(enc_only=False) for g_mode enhanced / spectrogram (trainer.py:212-213) against the oracle."""
    from zs_b200.patchers import Enhanced_Generator, Spectrogram_Patcher
    enc, dec, enc_sd, dec_sd = full_models
    if g_mode == 'spectrogram':
        gen_sd = syn.patcher_state_dict(4, c_in=513, c_out=513, c_h=1024, c_a=2)
        gen = Spectrogram_Patcher(ns=0.01, c_in=513, c_h=1024, c_a=2, seg_len=128)          # trainer.py:79
    else:
        gen_sd = syn.enhanced_generator_state_dict(4)
        gen = Enhanced_Generator(ns=0.01, dp=0.5, enc_size=1024, emb_size=1024, seg_len=128, n_speakers=102)   # trainer.py:77
    gen.load_state_dict(gen_sd, strict=True)
    path = AutoencoderPath(enc, dec, gen, g_mode=g_mode)
    B, T = 2, 128
    x = syn.spectrogram_batch(B, T, 81)
    c = torch.tensor([100, 101])
    u = syn.gumbel_uniform((B, 16, 1024), 81)
    x_dec, enc_np = path.test_step(x.permute(0, 2, 1), c, enc_only=False, noise=gumbel_from_uniform(u))
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        act = torch.from_numpy(enc_np)
        o = orc.combine_generator(orc.decoder_forward(dec_sd, act, c), act, c, gen_sd, g_mode, 100)
    # 'enhanced' chains a second Encoder + Decoder behind the first Decoder: its input already carries the first path's 1e-3
    # error, which a random-init network amplifies - the stand-alone patchers are held to SPEC_RELRMS above
    assert relrms(torch.from_numpy(x_dec), o) < (3e-2 if g_mode == 'enhanced' else SPEC_RELRMS)
    with pytest.raises(RuntimeError, match='target speakers'):
        path.test_step(x.permute(0, 2, 1), torch.tensor([3, 101]), enc_only=False, noise=gumbel_from_uniform(u))


def test_encode_to_files_writes_the_reference_unit_format(full_models, tmp_path):
    """test_encode (convert.py:342-360): one unit file per utterance, byte-identical to write_encodings of encode()'s rows."""
    from zs_b200.frontend import encode_to_files, write_encodings
    enc, dec, _, _ = full_models
    path = AutoencoderPath(enc, dec, seg_len=128)
    rng = np.random.Generator(np.random.PCG64(17))
    named = [(f'utt{i}', np.clip(rng.random((L, 513), dtype=np.float32), 1e-8, 1)) for i, L in enumerate((5, 64, 300))]
    done = encode_to_files(path, named, str(tmp_path), noise_seed=5)
    assert done == ['utt0', 'utt1', 'utt2']
    rows = path.encode_utterances([s for _, s in named], noise_seed=5)
    for (name, _), r in zip(named, rows):
        write_encodings(str(tmp_path / 'ref.txt'), r)
        assert (tmp_path / (name + '.txt')).read_bytes() == (tmp_path / 'ref.txt').read_bytes()
        assert r.shape[1] == 1024 and (r.sum(1) == 1).all()


def test_cta_pair_gemm_is_bit_identical_to_single_cta(full_models):
    """Layers with an even number of channel tiles and of segments per tile run as CTA pairs (tcgen05 cta_group::2, M = 256, each CTA
    staging half of the tile's columns); `zs_set_gemm_pair_mode(0)` runs one CTA per tile everywhere.  Same MMAs on the same
    operands in the same order per output element: the two must agree bit for bit - whole path (every epilogue variant), ragged and
    short segments, batches that leave tiles partly empty, and single layers through the C ABI."""
    lib = _lib.lib()
    enc, dec, _, _ = full_models
    try:
        for B, T in ((6, 128), (37, 128), (4, 77), (2, 207), (5, 9), (40, 16), (300, 128)):
            x = syn.spectrogram_batch(B, T, 93).cuda()
            c = syn.speaker_ids(B, 102, 93).cuda()
            noise = gumbel_from_uniform(syn.gumbel_uniform((B, Encoder.t8(T), 1024), 93)).cuda()
            res = []
            for mode in (0x500, 0xa02):  # 0x500: one CTA per tile, three-stage ring, every tap re-stages its tile; 0xa02: pairs, tap reuse (in pairs too) and the four-stage ring wherever they apply
                lib.zs_set_gemm_pair_mode(mode)
                act, logits, ids = enc.encode(x, noise)
                spec = dec.decode(None, c, unit_ids=ids)
                torch.cuda.synchronize()
                res.append((logits.clone(), ids.clone(), spec.clone()))
            (l0, i0, s0), (l1, i1, s1) = res
            assert torch.equal(l0, l1) and torch.equal(i0, i1) and torch.equal(s0, s1), (B, T)
        for cs in (dict(B=6, C_in=513, C_out=256, T=77, k=3), dict(B=40, C_in=64, C_out=512, T=256, k=3), dict(B=34, C_in=96, C_out=1024, T=16, k=1)):
            torch.manual_seed(1)
            xx = torch.randn(cs['B'], cs['C_in'], cs['T'], device='cuda')
            W = torch.randn(cs['C_out'], cs['C_in'], cs['k'], device='cuda') / (cs['C_in'] * cs['k']) ** 0.5
            bb = torch.randn(cs['C_out'], device='cuda') * 0.1
            outs = []
            for mode in (0x500, 0xa02):
                lib.zs_set_gemm_pair_mode(mode)
                outs.append((gh.conv_cl_to_cl(xx, W, bb, lrelu=True, inorm=True, halo_out=2), gh.conv_cl(xx, W, bb, lrelu=True, act=1)))
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), cs
    finally:
        lib.zs_set_gemm_pair_mode(1)


def test_fp16_spectrogram_output(full_models):
    """Decoder.decode(out_dtype=float16): the fp32 result rounded once to fp16 (bit-exact to that rounding), also through the
    accumulate rule and the streaming front-end's fp16 host buffers."""
    from zs_b200.frontend import StreamingResynthesizer
    enc, dec, _, _ = full_models
    for B, T in ((3, 128), (2, 77)):
        x = syn.spectrogram_batch(B, T, 95).cuda()
        c = syn.speaker_ids(B, 102, 95).cuda()
        noise = gumbel_from_uniform(syn.gumbel_uniform((B, Encoder.t8(T), 1024), 95)).cuda()
        _, _, ids = enc.encode(x, noise)
        s32 = dec.decode(None, c, unit_ids=ids)
        s16 = dec.decode(None, c, unit_ids=ids, out_dtype=torch.float16)
        assert s16.dtype == torch.float16 and torch.equal(s16, s32.half())
        acc = s16.clone()
        dec.decode(None, c, unit_ids=ids, out=acc, accumulate=1)
        assert (acc.float() - 2 * s32).abs().max().item() < 2e-3
    S, T = 20, 128
    st = StreamingResynthesizer(enc, dec, micro_batch=8, n_buffers=2)
    xh = syn.spectrogram_batch(S, T, 96).pin_memory()
    ch = syn.speaker_ids(S, 102, 96).pin_memory()
    nz = gumbel_from_uniform(syn.gumbel_uniform((S, 16, 1024), 96)).pin_memory()
    out32, out16 = torch.zeros(S, 513, T).pin_memory(), torch.zeros(S, 513, T, dtype=torch.float16).pin_memory()
    st.run(xh, ch, out32, None, nz)
    torch.cuda.synchronize()
    st.run(xh, ch, out16, None, nz)
    torch.cuda.synchronize()
    assert torch.equal(out16, out32.half())
