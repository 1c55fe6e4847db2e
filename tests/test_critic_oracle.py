"""CPU: the critic / classifier restatement (oracle/critic_oracle.py) against the live-reference fixtures
(tests/golden/critic_*.npz, written by tests/golden/make_golden_critic.py from model/model.py:113-226)."""
import json
import os

import numpy as np
import pytest
import torch

import zs_b200  # noqa: F401
from zs_b200 import synthetic as syn
from oracle import critic_oracle as corc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CRITIC_CASES = ['critic_patch_b3', 'critic_patch_t64', 'critic_target_b2']


def load_critic_golden(name):
    z = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    m = json.loads(bytes(z['meta']).decode())
    x = syn.spectrogram_batch(m['B'], m['T'], 950 + m['seed'])
    sd = syn.critic_state_dict(m['seed'], n_class=m['n_class'], seg_len=m['T'], with_value=m['kind'] == 'patch')
    return m, x, sd, torch.from_numpy(z['val']), torch.from_numpy(z['logits'])


@pytest.mark.parametrize('name', CRITIC_CASES)
def test_critic_oracle_matches_the_live_reference(name):
    m, x, sd, val, logits = load_critic_golden(name)
    if m['kind'] == 'patch':
        v, lg = corc.patch_discriminator(sd, x, seg_len=m['T'])
        assert (v - val).abs().max().item() <= 2e-5
        assert torch.equal(corc.patch_discriminator(sd, x, seg_len=m['T'], classify=False), v)
    else:
        lg = corc.target_classifier(sd, x, seg_len=m['T'])
    assert lg.shape == logits.shape
    assert (lg - logits).abs().max().item() <= 2e-5


def test_critic_modules_keep_the_reference_checkpoint_contract():
    """Same parameter names and shapes as model/model.py:113-131 / 169-187 (strict load of the synthetic reference-layout
    state dicts the fixtures were made with); unsupported segment lengths raise like the reference."""
    from zs_b200 import critic as zc
    for cls, with_value, n_class in ((zc.PatchDiscriminator, True, 33), (zc.TargetClassifier, False, 2)):
        for seg_len in (128, 64, 32):
            net = cls(n_class=n_class, seg_len=seg_len)
            sd = syn.critic_state_dict(0, n_class=n_class, seg_len=seg_len, with_value=with_value)
            net.load_state_dict(sd, strict=True)
            assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    with pytest.raises(NotImplementedError):
        zc.PatchDiscriminator(seg_len=100)
    with pytest.raises(RuntimeError, match='CUDA'):
        zc.TargetClassifier().eval()(torch.zeros(1, 513, 128))
