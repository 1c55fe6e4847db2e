"""-m gpu parity of the 2-D critic / classifier forward (forward part; model/model.py:113-226) against the
live-reference fixtures tests/golden/critic_*.npz and the oracle restatement.

Tolerance (fp16 operands and fp16 activations through six InstanceNorm2d layers, fp32 accumulation; the reference computes
in fp32): max-abs <= 1e-2 on logits / values of magnitude ~1, rel-RMS <= 5e-3 (measured on B200: 2.0e-3 / 1.6e-3)."""
import pytest
import torch

import zs_b200  # noqa: F401
from zs_b200 import critic as zc, synthetic as syn
from oracle import critic_oracle as corc
from test_critic_oracle import CRITIC_CASES, load_critic_golden

pytestmark = pytest.mark.gpu

MAX_ABS, REL_RMS = 1e-2, 5e-3


def relrms(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize('name', CRITIC_CASES)
def test_critic_forward_matches_the_live_reference(name):
    m, x, sd, val, logits = load_critic_golden(name)
    if m['kind'] == 'patch':
        net = zc.PatchDiscriminator(n_class=m['n_class'], seg_len=m['T'])
    else:
        net = zc.TargetClassifier(n_class=m['n_class'], seg_len=m['T'])
    net.load_state_dict(sd, strict=True)
    net.cuda().eval()
    if m['kind'] == 'patch':
        v, lg = net(x.cuda(), classify=True)
        assert torch.equal(net(x.cuda()), v)             # classify=False returns the same value
        print(f'{name}: value max-abs {(v.cpu() - val).abs().max().item():.2e}')
        assert (v.cpu() - val).abs().max().item() <= MAX_ABS
    else:
        lg = net(x.cuda())
        out = zc.classify(net, x.permute(0, 2, 1).numpy())       # Trainer.classify (trainer.py:230-235)
        assert out.shape == tuple(logits.shape) and (torch.from_numpy(out) - lg.cpu()).abs().max().item() == 0
    lg = lg.cpu()
    print(f'{name}: logits max-abs {(lg - logits).abs().max().item():.2e}, rel-RMS {relrms(lg, logits):.2e}')
    assert lg.shape == logits.shape and torch.isfinite(lg).all()
    assert (lg - logits).abs().max().item() <= MAX_ABS
    assert relrms(lg, logits) <= REL_RMS


def test_critic_batch32_vs_oracle_and_refusals():
    """Config-4 batch (B = 32): every sample agrees with the oracle; train mode and CPU inputs are refused loudly."""
    net = zc.PatchDiscriminator(n_class=33, seg_len=128)
    sd = syn.critic_state_dict(3, n_class=33, seg_len=128)
    net.load_state_dict(sd, strict=True)
    net.cuda().eval()
    x = syn.spectrogram_batch(32, 128, 960)
    v, lg = net(x.cuda(), classify=True)
    v_o, lg_o = corc.patch_discriminator(sd, x[:4])
    assert (v[:4].cpu() - v_o).abs().max().item() <= MAX_ABS and (lg[:4].cpu() - lg_o).abs().max().item() <= MAX_ABS
    # a sample's result does not depend on its position in the batch (statistics are per sample)
    v2, lg2 = net(x[28:].cuda(), classify=True)
    assert (v2 - v[28:]).abs().max().item() <= 1e-3 and (lg2 - lg[28:]).abs().max().item() <= 1e-3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xc = x.cuda()
    net(xc)
    e0.record()
    for _ in range(5):
        net(xc)
    e1.record()
    torch.cuda.synchronize()
    print(f'PatchDiscriminator forward, 32 x 128 frames: {e0.elapsed_time(e1) / 5:.3f} ms')
    with pytest.raises(RuntimeError, match='CUDA'):
        net(x[:1])
    net.train()
    with pytest.raises(NotImplementedError, match='train-mode'):
        net(xc)
