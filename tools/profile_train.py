"""Two pretrain_AE steps at B = 32 (the command ncu wraps for the training launch list; the second step is listed)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200
from zs_b200 import synthetic as syn, train as zt
from zs_b200.model import Decoder, Encoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot')
dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
enc.cuda().train(); dec.cuda().train()
step = zt.PretrainAE(enc, dec)
x, c = syn.spectrogram_batch(B, 128, 0).cuda(), syn.speaker_ids(B, 102, 0).cuda()
for i in range(2):
    loss = step.step(x, c)
torch.cuda.synchronize()
print('ok', loss.item())
