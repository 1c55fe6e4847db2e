import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200
from zs_b200 import synthetic as syn
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot'); dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot')); dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
enc.cuda().eval(); dec.cuda().eval()
x = syn.spectrogram_batch(B, 128, 0).cuda(); c = syn.speaker_ids(B, 102, 0).cuda()
noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 0)).cuda()
out = torch.empty(B, 513, 128, device='cuda')
for _ in range(3):
    _, _, ids = enc.encode(x, noise); dec.decode(None, c, unit_ids=ids, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 20
for _ in range(N):
    _, _, ids = enc.encode(x, noise); dec.decode(None, c, unit_ids=ids, out=out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'B={B}: CPU issue time {1e3 * (t1 - t0) / N:.3f} ms per encode+decode, GPU-complete {1e3 * (t2 - t0) / N:.3f} ms')
# CUDA graph replay
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    _, _, ids = enc.encode(x, noise); dec.decode(None, c, unit_ids=ids, out=out)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        _, _, ids = enc.encode(x, noise); dec.decode(None, c, unit_ids=ids, out=out)
torch.cuda.synchronize()
ref = out.clone()
out.zero_()
g.replay(); torch.cuda.synchronize()
print('graph replay equal:', torch.equal(out, ref))
t0 = time.perf_counter()
for _ in range(N): g.replay()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'graph: CPU issue {1e3 * (t1 - t0) / N:.3f} ms, GPU-complete {1e3 * (t2 - t0) / N:.3f} ms per replay')
