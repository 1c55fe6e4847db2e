"""Golden fixtures of the 2-D critic / classifier forward passes from the LIVE reference (build container only):

    python tests/golden/make_golden_critic.py

Imports the reference's own `model/model.py`, loads `zs_b200.synthetic.critic_state_dict` weights into
`PatchDiscriminator` / `TargetClassifier` (model/model.py:113-226) with `strict=True`, runs them in eval mode on CPU fp32
and stores the outputs.  Inputs and weights are re-derived from their seeds by the tests.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import zs_b200  # noqa: E402,F401
from zs_b200 import synthetic as syn  # noqa: E402
from make_golden import _load_ref_model  # noqa: E402

CASES = [
    ('critic_patch_b3', dict(kind='patch', B=3, T=128, n_class=33, seed=0)),
    ('critic_patch_t64', dict(kind='patch', B=2, T=64, n_class=8, seed=1)),
    ('critic_target_b2', dict(kind='target', B=2, T=128, n_class=2, seed=2)),
]


def main():
    torch.set_num_threads(os.cpu_count())
    ref = _load_ref_model()
    for name, m in CASES:
        x = syn.spectrogram_batch(m['B'], m['T'], 950 + m['seed'])
        sd = syn.critic_state_dict(m['seed'], n_class=m['n_class'], seg_len=m['T'], with_value=m['kind'] == 'patch')
        with torch.no_grad():
            if m['kind'] == 'patch':
                net = ref.PatchDiscriminator(n_class=m['n_class'], seg_len=m['T']).eval()
                net.load_state_dict(sd, strict=True)
                val, logits = net(x, classify=True)
            else:
                net = ref.TargetClassifier(n_class=m['n_class'], seg_len=m['T']).eval()
                net.load_state_dict(sd, strict=True)
                logits = net(x)
                val = torch.zeros(m['B'])
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, val=val.numpy(), logits=logits.numpy(), meta=np.frombuffer(json.dumps(m).encode(), dtype=np.uint8))
        print(f'{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB', tuple(logits.shape), val.numpy())


if __name__ == '__main__':
    main()
