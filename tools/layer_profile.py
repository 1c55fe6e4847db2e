"""Warm per-launch table of one encode->decode micro-batch (CUDA events around every launch of the library, after
warm-up, inputs rotating): layer, ms, algorithmic TFLOP/s, share.  python tools/layer_profile.py [MB] [reps] [warm] [pair]
warm = untimed steps before the table (default 5: burst clocks; ~400 puts a B200 into its power-capped sustained regime first),
pair = 0 turns the CTA-pair GEMM off, 0x101 keeps pairs but turns the four-stage single-CTA ring off (zs_set_gemm_pair_mode) for A/B tables."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402
from zs_b200 import _lib, synthetic as syn  # noqa: E402
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform  # noqa: E402

NAMES = ['pack x (bank in + cat)', 'bank', 'conv2 IN', 'conv3', 'conv4 s2 IN+avg', 'conv5', 'conv6 s2 IN+avg', 'conv7',
         'conv8 s2 IN+avg', 'dense1', 'dense2 IN+res', 'dense3', 'dense4 IN+res', 'gx', 'GRU enc', 'linear', 'bottleneck',
         'unit gather', 'd.conv1 PS', 'd.conv2 IN+up2', 'd.conv3 PS', 'd.conv4 IN+up2', 'd.conv5 PS', 'd.conv6 IN+up2',
         'd.dense1', 'd.dense2 IN+res', 'd.dense3', 'd.dense4 IN+res', 'd.gx', 'GRU dec', 'd.dense5', 'd.linear']


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 224
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    pair = int(sys.argv[4], 0) if len(sys.argv) > 4 else 1
    enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
    enc.cuda().eval()
    dec.cuda().eval()
    xs = [syn.spectrogram_batch(B, 128, i).cuda() for i in range(3)]
    c = syn.speaker_ids(B, 102, 0).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 0)).cuda()
    lib = _lib.lib()
    lib.zs_set_gemm_pair_mode(pair)

    def step(i):
        act, logits, ids = enc.encode(xs[i % 3], noise)
        return dec.decode(None, c, unit_ids=ids)
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    MAX = 256
    acc = None
    for r in range(reps):
        lib.zs_profile_begin()
        step(r)
        ms, fl, cl = (C.c_double * MAX)(), (C.c_double * MAX)(), (C.c_int * MAX)()
        n = lib.zs_profile_detail(ms, fl, cl, MAX)
        tot3, f3, c3 = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
        lib.zs_profile_end(tot3, f3, c3)
        rows = [(ms[i], fl[i], cl[i]) for i in range(n)]
        acc = rows if acc is None else [(a[0] + b[0], a[1], a[2]) for a, b in zip(acc, rows)]
    total = sum(a[0] for a in acc) / reps
    print(f'MB={B} warm={warm} pair={pair}: {len(acc)} launches, {total:.3f} ms per micro-batch (sum of per-launch events), {B * 128 / total / 1e3:.2f} M frames/s')
    kinds = {0: 'gemm', 1: 'gru', 2: 'other'}
    for i, (t, f, k) in enumerate(acc):
        t /= reps
        name = NAMES[i] if len(acc) == len(NAMES) else f'launch {i}'
        tf = f / (t * 1e-3) / 1e12 if t > 0 and f > 0 else 0.0
        print(f'{i:3d} {name:18s} {kinds[k]:5s} {t * 1e3:8.1f} us  {tf:7.1f} TFLOP/s  {100 * t / total:5.1f} %')


if __name__ == '__main__':
    main()
