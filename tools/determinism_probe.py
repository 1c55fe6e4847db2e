import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200
from zs_b200 import synthetic as syn
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform
enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot'); dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot')); dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
enc.cuda().eval(); dec.cuda().eval()
for B in (32, 64, 120, 128, 256):
    x = syn.spectrogram_batch(B, 128, 0).cuda(); c = syn.speaker_ids(B, 102, 0).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 0)).cuda()
    res = []
    for rep in range(3):
        act, logits, ids = enc.encode(x, noise)
        spec = dec.decode(None, c, unit_ids=ids)
        torch.cuda.synchronize()
        res.append((logits.clone(), ids.clone(), spec.clone()))
    for rep in (1, 2):
        dl = (res[rep][0] - res[0][0]).abs().max().item()
        di = (res[rep][1] != res[0][1]).sum().item()
        ds = (res[rep][2] - res[0][2]).abs()
        bad = (ds.amax(dim=(1, 2)) > 0).nonzero().flatten().tolist()
        print(f'B={B} rep{rep}: logits maxdiff {dl:.3e}, ids differing {di}, spec maxdiff {ds.max().item():.3e}, bad segments {bad[:12]}{"..." if len(bad) > 12 else ""} ({len(bad)})')
    # against the B=32-by-32 result
    if B > 32:
        ok = True
        for s0 in range(0, B, 32):
            a, l, i = enc.encode(x[s0:s0 + 32].contiguous(), noise[s0:s0 + 32])
            sp = dec.decode(None, c[s0:s0 + 32], unit_ids=i)
            if not torch.equal(sp, res[0][2][s0:s0 + 32]):
                d = (sp - res[0][2][s0:s0 + 32]).abs().amax(dim=(1, 2))
                print(f'   vs 32-batches at {s0}: logits equal {torch.equal(l, res[0][0][s0:s0+32])}, spec maxdiff {d.max().item():.3e}, bad {(d > 0).nonzero().flatten().tolist()[:10]}')
                ok = False
        print('   matches 32-batch results:', ok)
