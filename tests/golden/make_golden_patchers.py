"""Golden fixtures of the alternate TTS patchers from the LIVE reference (build container only: needs /root/reference):

    python tests/golden/make_golden_patchers.py

Imports the reference's own `model/model.py`, loads `zs_b200.synthetic` weights into `Spectrogram_Patcher` and
`Enhanced_Generator` (model/model.py:492-552) with `strict=True`, runs them in eval mode on CPU fp32 on a synthetic
decoded spectrogram and stores the outputs.  Inputs and weights are re-derived from their seeds by the tests.
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


def main():
    torch.set_num_threads(os.cpu_count())
    ref = _load_ref_model()
    cases = [
        ('patcher_small', dict(kind='spectrogram', B=3, T=40, c_in=33, c_h=64, c_a=2, seed=0)),
        ('patcher_small_t207', dict(kind='spectrogram', B=1, T=207, c_in=33, c_h=64, c_a=3, seed=1)),
        ('patcher_full', dict(kind='spectrogram', B=1, T=128, c_in=513, c_h=1024, c_a=2, seed=0)),
        ('enhanced_small', dict(kind='enhanced', B=2, T=48, c_in=33, c_h=(16, 64, 16), enc_size=32, emb_size=64, n_spk=5, seed=0)),
        ('enhanced_full', dict(kind='enhanced', B=1, T=128, c_in=513, c_h=(128, 512, 128), enc_size=1024, emb_size=1024, n_spk=102, seed=0)),
    ]
    for name, m in cases:
        x = syn.spectrogram_batch(m['B'], m['T'], 900 + m['seed'], c_in=m['c_in'])
        with torch.no_grad():
            if m['kind'] == 'spectrogram':
                sd = syn.patcher_state_dict(m['seed'], c_in=m['c_in'], c_out=m['c_in'], c_h=m['c_h'], c_a=m['c_a'])
                net = ref.Spectrogram_Patcher(c_in=m['c_in'], c_out=m['c_in'], c_h=m['c_h'], c_a=m['c_a'], ns=0.01, seg_len=128).eval()
                net.load_state_dict(sd, strict=True)
                c = syn.speaker_ids(m['B'], m['c_a'], m['seed'])
            else:
                sd = syn.enhanced_generator_state_dict(m['seed'], c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2],
                                                       enc_size=m['enc_size'], emb_size=m['emb_size'], n_speakers=m['n_spk'])
                net = ref.Enhanced_Generator(ns=0.01, dp=0.5, enc_size=m['enc_size'], emb_size=m['emb_size'], seg_len=128, n_speakers=m['n_spk'])
                if m['c_in'] != 513:       # the reference hard-codes 513 bins / (128, 512, 128) widths: rebuild its parts narrow
                    net.Encoder = ref.Encoder(c_in=m['c_in'], c_h1=m['c_h'][0], c_h2=m['c_h'][1], c_h3=m['c_h'][2], ns=0.01, dp=0.5,
                                              enc_size=m['enc_size'], seg_len=128, enc_mode='continues')
                    net.Decoder = ref.Decoder(c_in=m['enc_size'], c_out=m['c_in'], c_h=m['emb_size'], c_a=m['n_spk'], ns=0.01, seg_len=128)
                net = net.eval()
                net.load_state_dict(sd, strict=True)
                c = syn.speaker_ids(m['B'], 2, m['seed'])        # the Trainer calls it with c - shift in {0, 1}
            y = net(x, c)
        meta = dict(m, c_h=list(m['c_h']) if isinstance(m['c_h'], tuple) else m['c_h'])
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, out=y.numpy(), c=c.numpy(), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f'{name}: wrote {os.path.getsize(path) / 1024:.0f} KiB', tuple(y.shape))


if __name__ == '__main__':
    main()
