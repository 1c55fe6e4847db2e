"""Pins the oracle (oracle/ae_oracle.py) against outputs of the live reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

import zs_b200  # noqa: F401
from zs_b200 import synthetic as syn
from oracle import ae_oracle as orc
from conftest import GOLDEN, GOLDEN_CASES, load_golden


def _weights(meta):
    enc_sd = syn.encoder_state_dict(meta['seed'], c_in=meta['c_in'], c_h1=meta['c_h'][0], c_h2=meta['c_h'][1],
                                    c_h3=meta['c_h'][2], enc_size=meta['enc_size'], enc_mode=meta['enc_mode'])
    dec_sd = syn.decoder_state_dict(meta['seed'], c_in=meta['enc_size'], c_out=meta['c_in'],
                                    c_h=meta['emb_size'], c_a=meta['n_spk'])
    return enc_sd, dec_sd


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    m = g['meta']
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd = _weights(m)
    x = syn.spectrogram_batch(m['B'], m['T'], m['seed'], c_in=m['c_in'])
    c = syn.speaker_ids(m['B'], m['n_spk'], m['seed'])
    u = torch.from_numpy(g['uniform']) if 'uniform' in g else None
    with torch.no_grad():
        act, logits, ids = orc.encoder_forward(enc_sd, x, u, ns=m['ns'], seg_len=m['seg_len'],
                                               enc_mode=m['enc_mode'])
        spec = orc.decoder_forward(dec_sd, act, c, ns=m['ns'], seg_len=m['seg_len'])
    assert logits.shape == g['logits'].shape
    # fp32 summation-order noise only; the 9-frame case normalises over 2-frame instances
    # (InstanceNorm at T/4, T/8), which amplifies that noise ~10x
    tol = 2e-4 if m['T'] < 16 else 2e-5
    np.testing.assert_allclose(logits.numpy(), g['logits'], atol=tol, rtol=0)
    if m['enc_mode'] == 'one_hot':
        # unit ids: bit-exact, and the direct argmax(l+g) the CUDA path uses agrees with softmax-then-max
        assert np.array_equal(ids.numpy().astype(np.int32), g['act_argmax'])
        direct = (logits.permute(0, 2, 1) + orc.gumbel_noise(u)).argmax(-1)
        assert np.array_equal(direct.numpy().astype(np.int32), g['act_argmax'])
        assert set(np.unique(act.numpy())) <= {0.0, 1.0}
    if 'act' in g:
        np.testing.assert_allclose(act.numpy(), g['act'].astype(np.float32), atol=2e-5, rtol=0)
    assert spec.shape == g['spec'].shape
    np.testing.assert_allclose(spec.numpy(), g['spec'], atol=tol, rtol=0)
    if m['patch']:
        gen_sd = syn.decoder_state_dict(m['seed'] + 7, c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=2)
        c_t = torch.from_numpy(g['c_target'])
        with torch.no_grad():
            patched, *_ = orc.test_step(enc_sd, dec_sd, x, c_t, u, ns=m['ns'], seg_len=m['seg_len'],
                                        enc_mode=m['enc_mode'], gen_sd=gen_sd, g_mode='targeted',
                                        shift=m['n_spk'] - 2)
        np.testing.assert_allclose(patched.numpy(), g['spec_patched'], atol=4e-5, rtol=0)


def test_segment_plan_matches_reference_encode():
    plans = json.load(open(os.path.join(GOLDEN, 'chunk_plans.json')))
    assert len(plans) > 500
    for key, ref in plans.items():
        seg_len, L = map(int, key.split(':'))
        if 'error' in ref:
            with pytest.raises(Exception):
                orc.segment_plan(L, seg_len)
            continue
        padded, plan, keep = orc.segment_plan(L, seg_len)
        assert [[s, e - s] for s, e in plan] == ref['calls'], key
        units = sum((((e - s + 1) // 2 + 1) // 2 + 1) // 2 for s, e in plan)
        if keep is not None:
            units = min(units, keep)
        assert units == ref['n_units'], key


def test_format_encodings():
    enc = np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0]], dtype=np.float32)
    assert orc.format_encodings(enc) == '0 1 0\n1 0 0\n'


# ---- pretrain_AE step (trainer.py:321-332): oracle autograd / clip / Adam vs the live reference ----------------
TRAIN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('train_') and f.endswith('.npz'))


def load_train_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    g['meta'] = json.loads(bytes(g['meta']).decode())
    return g


def train_inputs(g):
    """(enc_sd, dec_sd, x, c, uniform, keep_masks) of a training fixture, re-derived from its seeds + stored draws."""
    m = g['meta']
    enc_sd, dec_sd = _weights(m)
    x = syn.spectrogram_batch(m['B'], m['T'], m['seed'], c_in=m['c_in'])
    c = syn.speaker_ids(m['B'], m['n_spk'], m['seed'])
    keep = None
    if m['dp'] > 0:
        keep = []
        for i, shp in enumerate(orc.dropout_mask_shapes(m['B'], m['T'], m['c_h'][1])):
            n = int(np.prod(shp))
            keep.append(torch.from_numpy(np.unpackbits(g[f'keep{i}'])[:n].astype(np.float32)).view(shp))
    return enc_sd, dec_sd, x, c, torch.from_numpy(g['uniform']), keep


def sample_idx(numel, n=512):
    step = max(1, numel // n)
    return np.arange(0, numel, step)[:n]


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_oracle_train_step_matches_reference(name):
    g = load_train_golden(name)
    m = g['meta']
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd, x, c, u, keep = train_inputs(g)
    loss, ge, gd, _, ids = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, keep, m['dp'], m['ns'], m['seg_len'])
    assert abs(loss.item() - float(g['loss'])) < 1e-6
    assert np.array_equal(ids.numpy().astype(np.int32), g['ids'])
    for net, grads in (('enc', ge), ('dec', gd)):
        for k, gr in grads.items():
            ref_n = float(g[f'gn:{net}:{k}'])
            assert abs(gr.norm().item() - ref_n) <= 2e-3 * ref_n + 1e-9, (net, k)
            ref_s = g[f'g:{net}:{k}']
            got = gr.reshape(-1)[sample_idx(gr.numel())].numpy()
            assert np.abs(got - ref_s).max() <= 2e-3 * np.abs(ref_s).max() + 1e-10, (net, k)
    # per-network clipping + one Adam step over both networks
    n_enc, _ = orc.clip_grad_norm(ge, m['max_grad_norm'])
    n_dec, _ = orc.clip_grad_norm(gd, m['max_grad_norm'])
    assert abs(n_enc.item() - float(g['norm_enc'])) < 1e-3 * float(g['norm_enc'])
    assert abs(n_dec.item() - float(g['norm_dec'])) < 1e-3 * float(g['norm_dec'])
    state = {}
    orc.pretrain_ae_step(enc_sd, dec_sd, state, x, c, u, keep, m['dp'], m['ns'], m['seg_len'], lr=m['lr'],
                         max_grad_norm=m['max_grad_norm'])
    for net, sd in (('enc', enc_sd), ('dec', dec_sd)):
        for k, p in sd.items():
            got = p.reshape(-1)[sample_idx(p.numel())].numpy()
            # the first Adam step moves a weight by lr * g / (|g| + eps): where |g| is within ~100 eps of zero the
            # update direction is summation-order noise, so those elements only have to stay within one lr
            tol = np.where(np.abs(g[f'g:{net}:{k}']) > 1e-6, 2e-5, 1.1 * m['lr'])
            assert (np.abs(got - g[f'p:{net}:{k}']) <= tol).all(), f'{net}:{k}'


def test_torch_module_baseline_matches_the_oracle():
    """oracle/torch_modules.py (stock nn.Conv1d / nn.GRU / ... modules: bench.py's eager-CUDA bar) computes what the
    functional oracle computes, on the CPU, sharing one reference-layout state_dict (strict load)."""
    from oracle.torch_modules import TorchDecoder, TorchEncoder
    from zs_b200 import synthetic as syn
    for T in (77, 128):
        esd = syn.encoder_state_dict(3, enc_size=64, enc_mode='one_hot', c_h1=16, c_h2=64, c_h3=16, c_in=33)
        dsd = syn.decoder_state_dict(3, c_in=64, c_h=64, c_a=5, c_out=33)
        te = TorchEncoder(33, 16, 64, 16, 0.01, 64, 128).eval()
        td = TorchDecoder(64, 33, 64, 5, 0.01, 128).eval()
        te.load_state_dict(esd, strict=True)
        td.load_state_dict(dsd, strict=True)
        x, c = syn.spectrogram_batch(3, T, 1, c_in=33), syn.speaker_ids(3, 5, 1)
        u = syn.gumbel_uniform((3, (T + 7) // 8, 64), 1)
        with torch.no_grad():
            act, logits = te(x, orc.gumbel_noise(u))
            spec = td(act, c)
            o_act, o_logits, _ = orc.encoder_forward(esd, x, u)
            o_spec = orc.decoder_forward(dsd, o_act, c)
        assert (logits - o_logits).abs().max() < 1e-4 and torch.equal(act, o_act) and (spec - o_spec).abs().max() < 1e-5


PATCHER_CASES = ['patcher_small', 'patcher_small_t207', 'patcher_full', 'enhanced_small', 'enhanced_full']


def patcher_inputs(g):
    m = g['meta']
    x = syn.spectrogram_batch(m['B'], m['T'], 900 + m['seed'], c_in=m['c_in'])
    if m['kind'] == 'spectrogram':
        sd = syn.patcher_state_dict(m['seed'], c_in=m['c_in'], c_out=m['c_in'], c_h=m['c_h'], c_a=m['c_a'])
    else:
        sd = syn.enhanced_generator_state_dict(m['seed'], c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2],
                                               enc_size=m['enc_size'], emb_size=m['emb_size'], n_speakers=m['n_spk'])
    return sd, x, torch.from_numpy(g['c'])


@pytest.mark.parametrize('name', PATCHER_CASES)
def test_oracle_patchers_match_live_reference(name):
    """Spectrogram_Patcher / Enhanced_Generator (model/model.py:492-552): the oracle's restatement against outputs of the
    live reference modules (tests/golden/make_golden_patchers.py)."""
    g = load_golden(name)
    sd, x, c = patcher_inputs(g)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        y = (orc.spectrogram_patcher_forward(sd, x, c) if g['meta']['kind'] == 'spectrogram'
             else orc.enhanced_generator_forward(sd, x, c))
    assert y.shape == g['out'].shape
    assert np.abs(y.numpy() - g['out']).max() <= 2e-5
