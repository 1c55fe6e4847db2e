"""Generate the TRAINING golden fixtures (`train_*.npz`) from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_train.py

For each case the reference's own `Encoder`/`Decoder` (model/model.py, unmodified, CPU fp32, train mode) run one
iteration of `Trainer.train(mode='pretrain_AE')` (trainer.py:321-332) written out with the reference's own
helpers' semantics: encode_step -> decode_step -> `torch.mean(torch.abs(x_dec - x))` -> zero_grad -> backward ->
`nn.utils.clip_grad_norm_` per network (utils.py:53-55) -> `optim.Adam(lr=1e-4, betas=(0.5, 0.9))` step
(trainer.py:64-66).  RNG: `torch.manual_seed` right before the encoder call; on CPU in train mode the six Dropout
keep-masks are drawn from the CPU generator (bernoulli_ on a tensor shaped like the dropout input) BEFORE the
Gumbel `torch.rand`, so the script re-draws both in that order from the same seed and stores them; the replay
in tests/test_oracle_golden.py feeds them to the oracle explicitly.

Stored per case: the draws, the loss, the two pre-clip gradient norms, and for EVERY parameter tensor its
gradient L2 norm plus a strided sample (<= 512 values) of the gradient and of the parameter after the Adam step.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

import zs_b200  # noqa: E402,F401
from zs_b200 import synthetic as syn  # noqa: E402
from oracle import ae_oracle as orc  # noqa: E402

N_SAMPLE = 512


def sample_idx(numel):
    step = max(1, numel // N_SAMPLE)
    return np.arange(0, numel, step)[:N_SAMPLE]


def _load_ref_model():
    spec = importlib.util.spec_from_file_location('ref_model', os.path.join(REF, 'model', 'model.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_case(ref, name, *, B, T, enc_size, emb_size, n_spk, dp, c_h=(128, 512, 128), c_in=513, seg_len=128, ns=0.01,
             seed=0, lr=1e-4, max_grad_norm=5.0):
    enc_kw = dict(c_in=c_in, c_h1=c_h[0], c_h2=c_h[1], c_h3=c_h[2], enc_size=enc_size, enc_mode='one_hot')
    enc_sd = syn.encoder_state_dict(seed, **enc_kw)
    dec_sd = syn.decoder_state_dict(seed, c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=n_spk)
    enc = ref.Encoder(ns=ns, dp=dp, seg_len=seg_len, **enc_kw).train()
    dec = ref.Decoder(c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=n_spk, ns=ns, seg_len=seg_len).train()
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    x = syn.spectrogram_batch(B, T, seed, c_in=c_in)
    c = syn.speaker_ids(B, n_spk, seed)
    T8 = (((T + 1) // 2 + 1) // 2 + 1) // 2
    out = {}

    # the draws the reference will make, in its order (dropout masks only when p > 0: F.dropout skips the RNG at p = 0)
    torch.manual_seed(4321 + seed)
    keep = None
    if dp > 0:
        keep = [torch.empty(s).bernoulli_(1 - dp) for s in orc.dropout_mask_shapes(B, T, c_h[1])]
        for i, k in enumerate(keep):
            out[f'keep{i}'] = np.packbits(k.numpy().astype(np.uint8).reshape(-1))
    uniform = torch.rand(B, T8, enc_size)
    out['uniform'] = uniform.numpy()

    params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.Adam(params, lr=lr, betas=(0.5, 0.9))          # trainer.py:64-66
    torch.manual_seed(4321 + seed)
    enc_act, _ = enc(x)                                              # trainer.py:325
    x_dec = dec(enc_act, c)                                          # :326
    loss = torch.mean(torch.abs(x_dec - x))                          # :327
    enc.zero_grad(); dec.zero_grad()                                 # :328 reset_grad
    loss.backward()                                                  # :329
    grads = {('enc', k): p.grad.detach().clone() for k, p in enc.named_parameters()}
    grads.update({('dec', k): p.grad.detach().clone() for k, p in dec.named_parameters()})
    n_enc = torch.nn.utils.clip_grad_norm_(enc.parameters(), max_grad_norm)   # :330 grad_clip, per network
    n_dec = torch.nn.utils.clip_grad_norm_(dec.parameters(), max_grad_norm)
    opt.step()                                                       # :332

    out['loss'] = np.float32(loss.item())
    out['norm_enc'] = np.float32(float(n_enc))
    out['norm_dec'] = np.float32(float(n_dec))
    out['ids'] = enc_act.detach().argmax(dim=1).numpy().astype(np.int32)
    for net_name, net in (('enc', enc), ('dec', dec)):
        for k, p in net.named_parameters():
            g = grads[(net_name, k)].reshape(-1)
            idx = sample_idx(g.numel())
            out[f'g:{net_name}:{k}'] = g[idx].numpy()
            out[f'gn:{net_name}:{k}'] = np.float32(g.norm().item())
            out[f'p:{net_name}:{k}'] = p.detach().reshape(-1)[idx].numpy()

    # self-check of the oracle restatement at generation time
    l2, ge, gd, _, ids2 = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, uniform, keep, dp, ns, seg_len)
    worst = 0.0
    for k, g in list(ge.items()) + list(gd.items()):
        net_name = 'enc' if k in ge and g is ge[k] else 'dec'
        ref_g = grads[(net_name, k)]
        worst = max(worst, float((g - ref_g).norm() / (ref_g.norm() + 1e-12)))
    print(f'{name}: loss ref {loss.item():.6f} oracle {l2.item():.6f}; norms {float(n_enc):.4f} / {float(n_dec):.4f}; '
          f'worst per-tensor rel grad error of the oracle {worst:.2e}')

    meta = dict(B=B, T=T, enc_size=enc_size, emb_size=emb_size, n_spk=n_spk, dp=dp, c_h=list(c_h), c_in=c_in,
                seg_len=seg_len, ns=ns, seed=seed, lr=lr, max_grad_norm=max_grad_norm, enc_mode='one_hot')
    out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f'{name}.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: wrote {os.path.getsize(path) / 1024:.0f} KiB')


def main():
    torch.set_num_threads(os.cpu_count())
    ref = _load_ref_model()
    small = dict(emb_size=64, n_spk=5, c_h=(16, 64, 16), c_in=33, enc_size=32)
    run_case(ref, 'train_small_dp0', B=3, T=128, dp=0.0, **small)
    run_case(ref, 'train_small_dp5', B=3, T=128, dp=0.5, **small)
    run_case(ref, 'train_full_b2_dp0', B=2, T=128, dp=0.0, enc_size=1024, emb_size=1024, n_spk=102)
    run_case(ref, 'train_full_b2_dp5', B=2, T=128, dp=0.5, enc_size=1024, emb_size=1024, n_spk=102, seed=3)


if __name__ == '__main__':
    main()
