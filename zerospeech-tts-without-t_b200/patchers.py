"""Drop-in `Enhanced_Generator` / `Spectrogram_Patcher` (reference model/model.py:492-552): the alternate TTS patchers
of `Trainer(g_mode='enhanced' | 'spectrogram')` (trainer.py:76-79), applied to the decoded spectrogram as
`x_dec += Generator(x_dec, c - shift)` (trainer.py:212-213, 280-281).

Same constructor signatures, same parameter names and shapes as the reference (checkpoints load with
`load_state_dict(strict=True)`), forward passes on libzsae.so.  Inference only: the stage-2 adversarial training that
updates a patcher (trainer.py:467-560) is outside the autoencoder hot path - a train-mode forward with autograd raises.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .model import Decoder, Encoder, _Packed, _ptr, _stream


class Enhanced_Generator(nn.Module):
    """model/model.py:492-502: an `Encoder(enc_mode='continues')` feeding a `Decoder`, both taken from this package."""

    def __init__(self, ns, dp, enc_size, emb_size, seg_len, n_speakers):
        super().__init__()
        self.Encoder = Encoder(ns=ns, dp=dp, enc_size=enc_size, seg_len=seg_len, enc_mode='continues')
        self.Decoder = Decoder(ns=ns, c_in=enc_size, c_h=emb_size, c_a=n_speakers, seg_len=seg_len)

    @property
    def operand(self):
        return self.Encoder.operand

    @operand.setter
    def operand(self, v):
        self.Encoder.operand = self.Decoder.operand = v

    @torch.no_grad()
    def patch(self, x, c, out=None, accumulate=0):
        """Generator(x, c) with the combine rule fused into the last kernel (`out += y` for accumulate = 1)."""
        act, _, _ = self.Encoder.encode(x)                 # 'continues': no noise, act = leaky_relu(logits)
        return self.Decoder.decode(act, c, out=out, accumulate=accumulate)

    def forward(self, x, c):
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError('Enhanced_Generator: the stage-2 training of the patcher is outside the autoencoder hot path '
                                      '(call .eval() / torch.no_grad() for inference)')
        return self.patch(x, c)


class Spectrogram_Patcher(_Packed):
    """model/model.py:503-549."""

    def __init__(self, c_in=512, c_out=513, c_h=512, c_a=8, ns=0.2, seg_len=64):
        super().__init__()
        self.ns, self.seg_len = ns, seg_len
        self.c_in, self.c_out, self.c_h, self.c_a = c_in, c_out, c_h, c_a
        self.input_layer = nn.Linear(c_in, c_h)
        for j in range(1, 5):
            setattr(self, f'dense{j}', nn.Linear(c_h, c_h))
        self.RNN = nn.GRU(input_size=c_h, hidden_size=c_h // 2, num_layers=1, bidirectional=True)
        self.dense5 = nn.Linear(2 * c_h + c_h, c_h)
        self.linear = nn.Linear(c_h, c_out)
        self.emb1 = nn.Embedding(c_a, c_h)
        self.emb2 = nn.Embedding(c_a, c_h)

    def _free_eval(self):
        if self._handle is not None:
            _lib.lib().zs_patcher_free(self._handle)
            self._handle = None

    def _free(self):
        self._free_eval()

    def weight_table(self):
        get = dict(self.named_parameters()).__getitem__
        w = _lib.PatcherWeights()
        w.input_w, w.input_b = get('input_layer.weight').data_ptr(), get('input_layer.bias').data_ptr()
        for i in range(4):
            w.dense_w[i] = get(f'dense{i + 1}.weight').data_ptr()
            w.dense_b[i] = get(f'dense{i + 1}.bias').data_ptr()
        for i, sfx in enumerate(('', '_reverse')):
            w.gru_w_ih[i] = get('RNN.weight_ih_l0' + sfx).data_ptr()
            w.gru_w_hh[i] = get('RNN.weight_hh_l0' + sfx).data_ptr()
            w.gru_b_ih[i] = get('RNN.bias_ih_l0' + sfx).data_ptr()
            w.gru_b_hh[i] = get('RNN.bias_hh_l0' + sfx).data_ptr()
        w.dense5_w, w.dense5_b = get('dense5.weight').data_ptr(), get('dense5.bias').data_ptr()
        w.linear_w, w.linear_b = get('linear.weight').data_ptr(), get('linear.bias').data_ptr()
        w.emb[0], w.emb[1] = get('emb1.weight').data_ptr(), get('emb2.weight').data_ptr()
        return w

    def _pack(self, dev, train=False):
        if train:
            raise NotImplementedError('Spectrogram_Patcher: inference only')
        for p in self.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('Spectrogram_Patcher parameters must be contiguous float32')
        cfg = _lib.PatcherCfg(self.c_in, self.c_out, self.c_h, self.c_a, _lib.OPERANDS[self.operand], self.ns)
        w = self.weight_table()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().zs_patcher_pack(C.byref(cfg), C.byref(w), _stream(), C.byref(h)))
        return h

    @torch.no_grad()
    def patch(self, x, c, out=None, accumulate=0):
        """Generator(x, c): x (B, c_in, T) fp32, c (B,) speaker ids -> (B, c_out, T); `out`/`accumulate` as in
        Decoder.decode (`out` may be `x` itself: x_dec += Generator(x_dec, c), trainer.py:212-213)."""
        self._check_input(x, 'x')
        if x.dim() != 3 or x.shape[1] != self.c_in:
            raise RuntimeError(f'Spectrogram_Patcher: expected (B, {self.c_in}, T), got {tuple(x.shape)}')
        dev = x.device
        x = x.detach().contiguous().float()
        B, _, T = x.shape
        if not c.is_cuda and c.numel() and (int(c.min()) < 0 or int(c.max()) >= self.c_a):
            raise RuntimeError(f'Spectrogram_Patcher: speaker id outside [0, {self.c_a})')
        c = c.to(dev, torch.int64).contiguous().view(-1)
        if c.numel() != B:
            raise RuntimeError(f'Spectrogram_Patcher: {c.numel()} speaker ids for {B} segments')
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._ensure_packed(dev)
            if out is None:
                if accumulate:
                    raise RuntimeError('Spectrogram_Patcher: accumulate needs `out`')
                out = torch.empty(B, self.c_out, T, dtype=torch.float32, device=dev)
            elif tuple(out.shape) != (B, self.c_out, T) or out.dtype != torch.float32 or not out.is_contiguous():
                raise RuntimeError('Spectrogram_Patcher: `out` must be a contiguous float32 (B, c_out, T) tensor')
            ws = self._get_workspace(lib.zs_patcher_workspace_bytes(h, B, T), dev)
            _lib.check(lib.zs_patcher_forward(h, _ptr(x), _ptr(c), B, T, _ptr(out), accumulate, _ptr(ws), ws.numel(), _stream()))
        return out

    def forward(self, x, c):
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError('Spectrogram_Patcher: the stage-2 training of the patcher is outside the autoencoder hot path '
                                      '(call .eval() / torch.no_grad() for inference)')
        return self.patch(x, c)
