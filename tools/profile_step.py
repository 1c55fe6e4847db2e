"""One encode->decode micro-batch, repeated: the command ncu wraps (launch list / --set full captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402
from zs_b200 import synthetic as syn  # noqa: E402
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
    enc.cuda().eval()
    dec.cuda().eval()
    x = syn.spectrogram_batch(B, 128, 0).cuda()
    c = syn.speaker_ids(B, 102, 0).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 0)).cuda()
    for _ in range(reps):
        act, logits, ids = enc.encode(x, noise)
        spec = dec.decode(None, c, unit_ids=ids)
    torch.cuda.synchronize()
    print('ok', float(spec.mean()))


if __name__ == '__main__':
    main()
