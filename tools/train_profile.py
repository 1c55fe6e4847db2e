"""Warm per-launch table of one EAGER pretrain_AE iteration (CUDA events around every library launch).
python tools/train_profile.py [B]"""
import collections
import csv
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402
from zs_b200 import _lib, synthetic as syn, train as zt  # noqa: E402
from zs_b200.model import Decoder, Encoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot')
dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
enc.cuda().train(); dec.cuda().train()
step = zt.PretrainAE(enc, dec, use_graph=False)
xs = [syn.spectrogram_batch(B, 128, s).cuda() for s in range(4)]
cs = [syn.speaker_ids(B, 102, s).cuda() for s in range(4)]
for i in range(4):
    step.step(xs[i % 4], cs[i % 4])
torch.cuda.synchronize()
lib = _lib.lib()
MAX, reps, acc = 512, 5, None
for r in range(reps):
    lib.zs_profile_begin()
    step.step(xs[r % 4], cs[r % 4])
    ms, fl, cl = (C.c_double * MAX)(), (C.c_double * MAX)(), (C.c_int * MAX)()
    n = lib.zs_profile_detail(ms, fl, cl, MAX)
    rows = [(ms[i], fl[i], cl[i]) for i in range(n)]
    names = [lib.zs_profile_name(i).decode() for i in range(n)]
    t3, f3, c3 = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
    lib.zs_profile_end(t3, f3, c3)
    acc = rows if acc is None else [(a[0] + b[0], a[1], a[2]) for a, b in zip(acc, rows)]
total = sum(a[0] for a in acc) / reps
print(f'B={B}: {len(acc)} launches, {total:.3f} ms per iteration (sum of per-launch events)')
agg = collections.OrderedDict()
for i, (t, f, k) in enumerate(acc):
    nm = names[i] if names else ('gemm', 'gru', 'other')[k]
    a = agg.setdefault(nm, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t / reps; a[2] += f
for nm, (c, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = f / (t * 1e-3) / 1e12 if f > 0 else 0
    print(f'  {nm[:48]:48s} n={c:3d} {t * 1e3:8.1f} us {100 * t / total:5.1f}%  {tf:7.1f} TFLOP/s')
if True:
    print('top launches:')
    for i, (t, f, k) in sorted(enumerate(acc), key=lambda it: -it[1][0])[:25]:
        print(f'  #{i:3d} {names[i][:40]:40s} {t / reps * 1e3:8.1f} us  {f / max(t / reps, 1e-9) / 1e9:7.1f} TFLOP/s')
if len(sys.argv) > 2:
    print(f'all {sys.argv[2]} launches in order:')
    for i, (t, f, k) in enumerate(acc):
        if sys.argv[2] in names[i]:
            print(f'  #{i:3d} {names[i][:32]:32s} {t / reps * 1e3:8.1f} us  {f / 1e9:7.2f} GFLOP  {f / max(t / reps, 1e-9) / 1e9:7.1f} TFLOP/s')
