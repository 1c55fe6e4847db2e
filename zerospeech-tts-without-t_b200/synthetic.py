"""Deterministic synthetic weights and inputs for the autoencoder hot path.

There is no network for checkpoints or data, so every test, the smoke run and
the benchmark use random-init weights of the reference architecture and
synthetic spectrograms of the reference's value range.  Everything here is
generated with numpy's PCG64 (stable across numpy versions and machines) so the
golden fixtures made in the build container can be re-derived bit-for-bit on
the GPU box without shipping 221 MB of weights.

Tensor names/shapes follow the reference checkpoint contract
(/root/reference/model/model.py:284-315 Decoder.__init__, :369-414
Encoder.__init__).  The value scale follows PyTorch's
default init (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear/GRU,
N(0,1) for nn.Embedding) so activations have the same statistics as a
freshly constructed reference model.
"""
from collections import OrderedDict

import numpy as np
import torch

N_BANK = 7  # conv1s kernel sizes 1..7 (model/model.py:375-377)


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def encoder_shapes(c_in=513, c_h1=128, c_h2=512, c_h3=128, enc_size=1024, enc_mode='one_hot'):
    """Ordered {name: (shape, fan_in)} of the reference Encoder state_dict."""
    s = OrderedDict()
    for i in range(N_BANK):
        k = i + 1
        s[f'conv1s.{i}.weight'] = ((c_h1, c_in, k), c_in * k)
        s[f'conv1s.{i}.bias'] = ((c_h1,), c_in * k)
    s['conv2.weight'] = ((c_h2, N_BANK * c_h1 + c_in, 1), N_BANK * c_h1 + c_in)
    s['conv2.bias'] = ((c_h2,), N_BANK * c_h1 + c_in)
    for j in range(3, 9):
        s[f'conv{j}.weight'] = ((c_h2, c_h2, 5), c_h2 * 5)
        s[f'conv{j}.bias'] = ((c_h2,), c_h2 * 5)
    for j in range(1, 5):
        s[f'dense{j}.weight'] = ((c_h2, c_h2), c_h2)
        s[f'dense{j}.bias'] = ((c_h2,), c_h2)
    for sfx in ('', '_reverse'):
        s[f'RNN.weight_ih_l0{sfx}'] = ((3 * c_h3, c_h2), c_h3)
        s[f'RNN.weight_hh_l0{sfx}'] = ((3 * c_h3, c_h3), c_h3)
        s[f'RNN.bias_ih_l0{sfx}'] = ((3 * c_h3,), c_h3)
        s[f'RNN.bias_hh_l0{sfx}'] = ((3 * c_h3,), c_h3)
    if enc_mode == 'multilabel_binary':
        n_out = 2 * enc_size
    elif enc_mode == 'binary':
        n_out = enc_size * enc_size
    else:
        n_out = enc_size
    s['linear.weight'] = ((n_out, c_h2 + 2 * c_h3), c_h2 + 2 * c_h3)
    s['linear.bias'] = ((n_out,), c_h2 + 2 * c_h3)
    return s


def decoder_shapes(c_in=1024, c_out=513, c_h=1024, c_a=102):
    """Ordered {name: (shape, fan_in)} of the reference Decoder state_dict (fan_in 0 = N(0,1))."""
    s = OrderedDict()
    for j in range(1, 7):
        co = 2 * c_h if j % 2 == 1 else c_h
        s[f'conv{j}.weight'] = ((co, c_h, 3), c_h * 3)
        s[f'conv{j}.bias'] = ((co,), c_h * 3)
    for j in range(1, 5):
        s[f'dense{j}.weight'] = ((c_h, c_h), c_h)
        s[f'dense{j}.bias'] = ((c_h,), c_h)
    hh = c_h // 2
    for sfx in ('', '_reverse'):
        s[f'RNN.weight_ih_l0{sfx}'] = ((3 * hh, c_h), hh)
        s[f'RNN.weight_hh_l0{sfx}'] = ((3 * hh, hh), hh)
        s[f'RNN.bias_ih_l0{sfx}'] = ((3 * hh,), hh)
        s[f'RNN.bias_hh_l0{sfx}'] = ((3 * hh,), hh)
    s['dense5.weight'] = ((c_h, 3 * c_h), 3 * c_h)
    s['dense5.bias'] = ((c_h,), 3 * c_h)
    s['linear.weight'] = ((c_out, c_h), c_h)
    s['linear.bias'] = ((c_out,), c_h)
    s['input_emb.weight'] = ((c_h, c_in), c_in)
    s['input_emb.bias'] = ((c_h,), c_in)
    for j in range(1, 6):
        s[f'emb{j}.weight'] = ((c_a, c_h), 0)
    return s


def patcher_shapes(c_in=513, c_out=513, c_h=1024, c_a=2):
    """Ordered {name: (shape, fan_in)} of the reference Spectrogram_Patcher state_dict (model/model.py:503-523)."""
    s = OrderedDict()
    s['input_layer.weight'] = ((c_h, c_in), c_in)
    s['input_layer.bias'] = ((c_h,), c_in)
    for j in range(1, 5):
        s[f'dense{j}.weight'] = ((c_h, c_h), c_h)
        s[f'dense{j}.bias'] = ((c_h,), c_h)
    hh = c_h // 2
    for sfx in ('', '_reverse'):
        s[f'RNN.weight_ih_l0{sfx}'] = ((3 * hh, c_h), hh)
        s[f'RNN.weight_hh_l0{sfx}'] = ((3 * hh, hh), hh)
        s[f'RNN.bias_ih_l0{sfx}'] = ((3 * hh,), hh)
        s[f'RNN.bias_hh_l0{sfx}'] = ((3 * hh,), hh)
    s['dense5.weight'] = ((c_h, 3 * c_h), 3 * c_h)
    s['dense5.bias'] = ((c_h,), 3 * c_h)
    s['linear.weight'] = ((c_out, c_h), c_h)
    s['linear.bias'] = ((c_out,), c_h)
    for j in (1, 2):
        s[f'emb{j}.weight'] = ((c_a, c_h), 0)
    return s


def _fill(shapes, seed):
    rng = _rng(seed)
    sd = OrderedDict()
    for name, (shape, fan_in) in shapes.items():
        if fan_in == 0:
            sd[name] = torch.from_numpy(rng.standard_normal(size=shape).astype(np.float32))
        else:
            sd[name] = _uniform(rng, shape, 1.0 / np.sqrt(fan_in))
    return sd


def encoder_state_dict(seed=0, **kw):
    return _fill(encoder_shapes(**kw), 1000 + seed)


def decoder_state_dict(seed=0, **kw):
    return _fill(decoder_shapes(**kw), 2000 + seed)


def patcher_state_dict(seed=0, **kw):
    return _fill(patcher_shapes(**kw), 6000 + seed)


def critic_shapes(n_class=33, seg_len=128, with_value=True):
    """PatchDiscriminator (with_value) / TargetClassifier (model/model.py:113-131, 169-187): name -> (shape, fan_in)."""
    kw = {128: 4, 64: 2, 32: 1}[seg_len]
    chans = [(1, 64), (64, 128), (128, 256), (256, 512), (512, 512)]
    sh = OrderedDict()
    for i, (ci, co) in enumerate(chans, 1):
        sh[f'conv{i}.weight'] = ((co, ci, 5, 5), ci * 25)
        sh[f'conv{i}.bias'] = ((co,), ci * 25)
    sh['conv6.weight'] = ((32, 512, 1, 1), 512)
    sh['conv6.bias'] = ((32,), 512)
    if with_value:
        sh['conv7.weight'] = ((1, 32, 17, kw), 32 * 17 * kw)
        sh['conv7.bias'] = ((1,), 32 * 17 * kw)
    sh['conv_classify.weight'] = ((n_class, 32, 17, kw), 32 * 17 * kw)
    sh['conv_classify.bias'] = ((n_class,), 32 * 17 * kw)
    if not with_value:      # TargetClassifier declares conv7 too (model/model.py:180-187) although its forward never uses it
        sh['conv7.weight'] = ((1, 32, 17, kw), 32 * 17 * kw)
        sh['conv7.bias'] = ((1,), 32 * 17 * kw)
    return sh


def critic_state_dict(seed=0, **kw):
    return _fill(critic_shapes(**kw), 7000 + seed)


def enhanced_generator_state_dict(seed=0, c_in=513, c_h1=128, c_h2=512, c_h3=128, enc_size=1024, emb_size=1024, n_speakers=102):
    """Enhanced_Generator (model/model.py:492-502): `Encoder.*` (continues mode) + `Decoder.*`."""
    sd = OrderedDict()
    for k, v in encoder_state_dict(seed + 50, c_in=c_in, c_h1=c_h1, c_h2=c_h2, c_h3=c_h3, enc_size=enc_size, enc_mode='continues').items():
        sd['Encoder.' + k] = v
    for k, v in decoder_state_dict(seed + 50, c_in=enc_size, c_out=c_in, c_h=emb_size, c_a=n_speakers).items():
        sd['Decoder.' + k] = v
    return sd


def spectrogram_batch(n_seg, n_frames, seed=0, c_in=513):
    """(B, c_in, T) float32 in [1e-8, 1] - the range preprocess.py:252 produces,
    already permuted the way trainer.py:196 hands it to the Encoder."""
    rng = _rng(3000 + seed)
    x = rng.random(size=(n_seg, n_frames, c_in), dtype=np.float32)
    x = np.clip(x, 1e-8, 1.0)
    return torch.from_numpy(x).permute(0, 2, 1).contiguous()


def speaker_ids(n_seg, n_speakers=102, seed=0):
    rng = _rng(4000 + seed)
    return torch.from_numpy(rng.integers(0, n_speakers, size=(n_seg,), dtype=np.int64))


def gumbel_uniform(shape, seed=0):
    """The uniform draw gumbel_softmax makes (model/model.py:96); the reference
    draws it from torch's CPU generator, tests pin it by passing it explicitly."""
    rng = _rng(5000 + seed)
    return torch.from_numpy(rng.random(size=tuple(shape), dtype=np.float32))
