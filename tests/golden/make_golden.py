"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own `model/model.py` (unmodified, by path), loads the
deterministic synthetic weights of `zs_b200.synthetic` into the reference
`Encoder`/`Decoder`, pins the Gumbel draw by seeding torch's CPU generator
immediately before each encoder call (model/model.py:96 draws `torch.rand` from
it), runs CPU fp32, and stores inputs that cannot be re-derived (the uniform
draw) plus the reference outputs.  Weights and spectrograms are re-derived from
their seeds by the tests.

It also replays the reference's chunking loop (`convert.py:183-221 encode()`)
with a recording stub trainer, after stubbing the audio/ASR imports convert.py
makes at module level (h5py, librosa, soundfile, speech_recognition, jiwer,
tensorboardX are absent from the image and unused by `encode()` itself).
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

import zs_b200  # noqa: E402
from zs_b200 import synthetic as syn  # noqa: E402


def _load_ref_model():
    spec = importlib.util.spec_from_file_location('ref_model', os.path.join(REF, 'model', 'model.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run_case(ref, name, *, B, T, enc_size, enc_mode, emb_size, n_spk, c_h=(128, 512, 128), c_in=513,
              seg_len=128, ns=0.01, patch=False, seed=0):
    enc_kw = dict(c_in=c_in, c_h1=c_h[0], c_h2=c_h[1], c_h3=c_h[2], enc_size=enc_size, enc_mode=enc_mode)
    enc_sd = syn.encoder_state_dict(seed, **enc_kw)
    dec_sd = syn.decoder_state_dict(seed, c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=n_spk)
    enc = ref.Encoder(ns=ns, dp=0.5, seg_len=seg_len, **enc_kw).eval()
    dec = ref.Decoder(c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=n_spk, ns=ns, seg_len=seg_len).eval()
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    x = syn.spectrogram_batch(B, T, seed, c_in=c_in)
    c = syn.speaker_ids(B, n_spk, seed)
    T8 = (((T + 1) // 2 + 1) // 2 + 1) // 2
    ushape = {'one_hot': (B, T8, enc_size), 'multilabel_binary': (B, T8, enc_size, 2),
              'gumbel_t': (B, enc_size, T8), 'binary': (B, T8, enc_size, enc_size), 'continues': None}[enc_mode]
    out = {}
    with torch.no_grad():
        torch.manual_seed(1234 + seed)
        if ushape is not None:
            out['uniform'] = torch.rand(ushape).numpy()
        torch.manual_seed(1234 + seed)          # same generator state -> same draw inside the reference
        act, logits = enc(x)
        spec = dec(act, c)
        out['logits'] = logits.numpy()
        out['act_argmax'] = act.argmax(dim=1).numpy().astype(np.int32)
        if enc_mode in ('multilabel_binary', 'gumbel_t', 'binary'):
            out['act'] = act.numpy().astype(np.uint8)
        if enc_mode == 'continues':
            out['act'] = act.numpy()
        out['spec'] = spec.numpy()
        if patch:   # trainer.py:208-209, g_mode='targeted', c in {n_spk-2, n_spk-1}
            gen_sd = syn.decoder_state_dict(seed + 7, c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=2)
            gen = ref.Decoder(c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=2, ns=ns, seg_len=seg_len).eval()
            gen.load_state_dict(gen_sd, strict=True)
            c_t = (c % 2) + (n_spk - 2)
            out['c_target'] = c_t.numpy()
            out['spec_patched'] = (dec(act, c_t) + gen(act, c_t - (n_spk - 2))).numpy()
    meta = dict(B=B, T=T, enc_size=enc_size, enc_mode=enc_mode, emb_size=emb_size, n_spk=n_spk,
                c_h=list(c_h), c_in=c_in, seg_len=seg_len, ns=ns, patch=patch, seed=seed)
    out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f'{name}.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: wrote {os.path.getsize(path) / 1024:.0f} KiB', {k: v.shape for k, v in out.items() if k != "meta"})


def _chunk_plans():
    """Replay convert.py:encode() over many lengths with a recording stub trainer."""
    for m in ('h5py', 'librosa', 'soundfile', 'speech_recognition', 'jiwer', 'tensorboardX', 'librosa.display',
              'unidecode', 'inflect', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['unidecode'].unidecode = lambda s: s
    sys.modules['inflect'].engine = lambda: None
    sys.modules['tensorboardX'].SummaryWriter = object
    sys.modules['jiwer'].wer = lambda *a, **k: 0.0
    sys.modules['scipy.signal'] = __import__('scipy.signal').signal
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import convert as ref_convert
    finally:
        os.chdir(cwd)

    class Recorder:
        def __init__(self):
            self.calls = []

        def encoder_test_step(self, tensor):
            t = tensor.shape[1]
            self.calls.append((int(tensor[0, 0, 0]), t))
            t8 = (((t + 1) // 2 + 1) // 2 + 1) // 2
            return np.zeros((1, 4, t8), dtype=np.float32)

    plans = {}
    for seg_len in (64, 128):
        for L in list(range(1, 3 * seg_len + 3)) + [777, 2000]:
            spec = np.zeros((L, 3), dtype=np.float32)
            spec[:, 0] = np.arange(L)          # frame index rides in channel 0
            rec = Recorder()
            try:
                enc = ref_convert.encode(spec, rec, seg_len, save=False)
                plans[f'{seg_len}:{L}'] = dict(calls=rec.calls, n_units=int(enc.shape[0]))
            except Exception as e:     # the reference raises on some lengths; record that too
                plans[f'{seg_len}:{L}'] = dict(error=type(e).__name__)
    with open(os.path.join(HERE, 'chunk_plans.json'), 'w') as f:
        json.dump(plans, f, separators=(',', ':'))
    print('chunk_plans.json:', len(plans), 'lengths')


def main():
    torch.set_num_threads(os.cpu_count())
    ref = _load_ref_model()
    small = dict(emb_size=64, n_spk=5, c_h=(16, 64, 16), c_in=33)
    if len(sys.argv) > 1 and sys.argv[1] == 'binary':      # added later: only this fixture (the others stay byte-identical)
        _run_case(ref, 'small_binary', B=2, T=48, enc_size=16, enc_mode='binary', **small)
        return
    _run_case(ref, 'small_onehot', B=3, T=40, enc_size=32, enc_mode='one_hot', patch=True, **small)
    _run_case(ref, 'small_onehot_odd', B=2, T=77, enc_size=32, enc_mode='one_hot', **small)
    _run_case(ref, 'small_mbv', B=2, T=48, enc_size=32, enc_mode='multilabel_binary', **small)
    _run_case(ref, 'small_continues', B=2, T=48, enc_size=32, enc_mode='continues', **small)
    _run_case(ref, 'small_gumbel_t', B=2, T=64, enc_size=32, enc_mode='gumbel_t', **small)
    _run_case(ref, 'small_binary', B=2, T=48, enc_size=16, enc_mode='binary', **small)
    _run_case(ref, 'small_zeropad', B=2, T=40, enc_size=32, enc_mode='one_hot', seg_len=32, **small)
    full = dict(emb_size=1024, n_spk=102)
    _run_case(ref, 'full_b2_t128', B=2, T=128, enc_size=1024, enc_mode='one_hot', patch=True, **full)
    _run_case(ref, 'full_b1_t207', B=1, T=207, enc_size=1024, enc_mode='one_hot', **full)
    _run_case(ref, 'full_b1_t9', B=1, T=9, enc_size=1024, enc_mode='one_hot', **full)
    _run_case(ref, 'full_b1_mbv', B=1, T=128, enc_size=1024, enc_mode='multilabel_binary', **full)
    _run_case(ref, 'full_b1_e512', B=1, T=128, enc_size=512, enc_mode='one_hot', **full)
    _chunk_plans()


if __name__ == '__main__':
    main()
