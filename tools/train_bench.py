"""pretrain_AE step timing (config 4: B = 32 per rank, seg_len 128): ms/step + per-kernel-class breakdown."""
import ctypes as C, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200
from zs_b200 import _lib, synthetic as syn, train as zt
from zs_b200.model import Decoder, Encoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot')
dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
enc.cuda().train(); dec.cuda().train()
step = zt.PretrainAE(enc, dec, use_graph=(len(sys.argv) <= 3 or sys.argv[3] != 'eager'),
                     async_wgrad=(len(sys.argv) <= 4 or sys.argv[4] != 'sync'))     # argv: B steps eager|graph sync|async
xs = [syn.spectrogram_batch(B, 128, s).cuda() for s in range(4)]
cs = [syn.speaker_ids(B, 102, s).cuda() for s in range(4)]
for i in range(5):
    loss = step.step(xs[i % 4], cs[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    loss = step.step(xs[i % 4], cs[i % 4])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f'B={B}: {ms:.3f} ms/step, {B * 128 / ms * 1e3:.0f} frames/s, loss {loss.item():.4f}, skipped {step.n_skipped}, scale {step.loss_scale}')
lib = _lib.lib()
lib.zs_profile_begin()
for i in range(3):
    step.step(xs[i % 4], cs[i % 4])
m, f, l = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
lib.zs_profile_end(m, f, l)
for i, n in enumerate(('gemm', 'gru', 'other')):
    print(f'  {n}: {m[i] / 3:.3f} ms/step, {l[i] // 3} launches, {f[i] / 3 / 1e9:.1f} GFLOP -> {f[i] / max(m[i], 1e-9) / 1e9:.1f} TFLOP/s')
t0 = time.perf_counter()
for i in range(steps):
    step.step(xs[i % 4], cs[i % 4])
t_issue = (time.perf_counter() - t0) / steps * 1e3
torch.cuda.synchronize()
print(f'  host issue time {t_issue:.3f} ms/step')
