"""Drop-in `Encoder` / `Decoder` for the reference's model/model.py:283-489.

Same constructor signatures, same parameter names and shapes (so a reference
`.pth` loads with `load_state_dict(strict=True)` and `state_dict()` round-trips,
SURVEY.md Appendix B), same `forward` contracts:

    Encoder.forward(x (B, c_in, T))            -> (out_act (B, enc_size, T8), out (B, n_out, T8))
    Decoder.forward(x (B, c_in, T8), c (B,))   -> (B, c_out, 8*T8)

but the forward passes run libzsae.so (hand-written sm_100a kernels, include/zs_ae.h)
instead of ATen/cuDNN.  The nn.Conv1d / nn.Linear / nn.GRU / nn.Embedding children are
parameter containers only (they fix the checkpoint layout and the default init);
they are never called.

Differences from the reference, all deliberate:
  * the Gumbel noise of the discrete bottleneck is an explicit, optional argument
    (`noise=`): the reference draws it inside forward from torch's CPU generator
    (model/model.py:96) - when `noise` is None this module draws it the same way
    (same generator, same shape, same call order) so seeding reproduces the reference;
  * inference only: forward does not build an autograd graph (`torch.no_grad`
    semantics); dropout (train mode) is not applied;
  * segments are limited to 9 <= T <= 256 frames (the range convert.py's chunking produces
    for seg_len = 128) and the reflect padding mode (hps seg_len >= 64);
  * errors are raised, never swallowed; there is no CPU path.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

GUMBEL_EPS = 1e-20


def sample_gumbel(shape):
    """model/model.py:95-98 on the CPU generator (the reference's RNG contract)."""
    u = torch.rand(shape)
    return -torch.log(-torch.log(u + GUMBEL_EPS) + GUMBEL_EPS)


def gumbel_from_uniform(u):
    return -torch.log(-torch.log(u + GUMBEL_EPS) + GUMBEL_EPS)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Packed(nn.Module):
    """Shared plumbing: lazy packing of the fp32 parameters into the library handle."""

    def __init__(self):
        super().__init__()
        self._handle = None
        self._packed_key = None
        self._workspace = None
        self.operand = 'fp16'

    def _param_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters()) + (self.operand,)

    def _free(self):
        raise NotImplementedError

    def _pack(self, dev):
        raise NotImplementedError

    def _ensure_packed(self, dev):
        key = self._param_key()
        if self._handle is None or key != self._packed_key:
            self._free()
            self._pack(dev)
            self._packed_key = key
        return self._handle

    def _get_workspace(self, nbytes, dev):
        ws = self._workspace
        if ws is None or ws.numel() < nbytes or ws.device != dev:
            self._workspace = ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return ws

    def _check_input(self, x, name):
        if not x.is_cuda:
            raise RuntimeError(f'{type(self).__name__}.forward: `{name}` must be a CUDA tensor - '
                               'this path has no CPU fallback')
        for p in self.parameters():
            if p.device != x.device:
                raise RuntimeError(f'{type(self).__name__}: parameters are on {p.device}, input on {x.device}')
            break

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


class Encoder(_Packed):
    def __init__(self, c_in=513, c_h1=128, c_h2=512, c_h3=128, ns=0.2, dp=0.5, enc_size=512, seg_len=64,
                 enc_mode='continues'):
        super().__init__()
        self.ns, self.dp, self.enc_size, self.seg_len, self.enc_mode = ns, dp, enc_size, seg_len, enc_mode
        self.c_in, self.c_h1, self.c_h2, self.c_h3 = c_in, c_h1, c_h2, c_h3
        self.conv1s = nn.ModuleList([nn.Conv1d(c_in, c_h1, kernel_size=k) for k in range(1, 8)])
        self.conv2 = nn.Conv1d(len(self.conv1s) * c_h1 + c_in, c_h2, kernel_size=1)
        for j in range(3, 9):
            setattr(self, f'conv{j}', nn.Conv1d(c_h2, c_h2, kernel_size=5, stride=2 if j % 2 == 0 else 1))
        for j in range(1, 5):
            setattr(self, f'dense{j}', nn.Linear(c_h2, c_h2))
        self.RNN = nn.GRU(input_size=c_h2, hidden_size=c_h3, num_layers=1, bidirectional=True)
        if enc_mode == 'multilabel_binary':
            n_out = 2 * enc_size
        elif enc_mode in ('continues', 'one_hot', 'gumbel_t'):
            assert enc_size % 2 == 0
            n_out = enc_size
        elif enc_mode == 'binary':
            raise NotImplementedError("enc_mode 'binary' (enc_size^2 projection) is not implemented on the B200 path")
        else:
            raise NotImplementedError('Invalid encoding mode!')
        self.n_out = n_out
        self.linear = nn.Linear(c_h2 + 2 * c_h3, n_out)

    # -- library handle ------------------------------------------------------------------
    def _free(self):
        if self._handle is not None:
            _lib.lib().zs_encoder_free(self._handle)
            self._handle = None

    def _pack(self, dev):
        lib = _lib.lib()
        cfg = _lib.EncoderCfg(self.c_in, self.c_h1, self.c_h2, self.c_h3, self.enc_size,
                              _lib.ENC_MODES[self.enc_mode], self.seg_len, _lib.OPERANDS[self.operand], self.ns)
        w = _lib.EncoderWeights()
        for p in self.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('Encoder parameters must be contiguous float32')
        for i in range(7):
            w.conv1s_w[i] = self.conv1s[i].weight.data_ptr()
            w.conv1s_b[i] = self.conv1s[i].bias.data_ptr()
            conv = getattr(self, f'conv{i + 2}')
            w.conv_w[i] = conv.weight.data_ptr()
            w.conv_b[i] = conv.bias.data_ptr()
        for i in range(4):
            d = getattr(self, f'dense{i + 1}')
            w.dense_w[i] = d.weight.data_ptr()
            w.dense_b[i] = d.bias.data_ptr()
        for i, sfx in enumerate(('', '_reverse')):
            w.gru_w_ih[i] = getattr(self.RNN, 'weight_ih_l0' + sfx).data_ptr()
            w.gru_w_hh[i] = getattr(self.RNN, 'weight_hh_l0' + sfx).data_ptr()
            w.gru_b_ih[i] = getattr(self.RNN, 'bias_ih_l0' + sfx).data_ptr()
            w.gru_b_hh[i] = getattr(self.RNN, 'bias_hh_l0' + sfx).data_ptr()
        w.linear_w = self.linear.weight.data_ptr()
        w.linear_b = self.linear.bias.data_ptr()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.zs_encoder_pack(C.byref(cfg), C.byref(w), _stream(), C.byref(h)))
        self._handle = h

    @staticmethod
    def t8(T):
        return (((T + 1) // 2 + 1) // 2 + 1) // 2

    def noise_shape(self, B, T):
        T8 = self.t8(T)
        return {'one_hot': (B, T8, self.enc_size), 'multilabel_binary': (B, T8, self.enc_size, 2),
                'gumbel_t': (B, self.enc_size, T8), 'continues': None}[self.enc_mode]

    @torch.no_grad()
    def encode(self, x, noise=None):
        """forward() plus the unit ids: returns (out_act, out, unit_ids or None)."""
        self._check_input(x, 'x')
        if x.dim() != 3 or x.shape[1] != self.c_in:
            raise RuntimeError(f'Encoder: expected (B, {self.c_in}, T), got {tuple(x.shape)}')
        B, _, T = x.shape
        dev = x.device
        x = x.detach().contiguous().float()
        shape = self.noise_shape(B, T)
        if shape is not None:
            if noise is None:
                noise = sample_gumbel(shape)            # CPU generator, like the reference
            if tuple(noise.shape) != shape:
                raise RuntimeError(f'Encoder: noise must have shape {shape}, got {tuple(noise.shape)}')
            noise = noise.to(dev, torch.float32, non_blocking=True).contiguous()
        else:
            noise = None
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._ensure_packed(dev)
            T8 = self.t8(T)
            logits = torch.empty(B, self.n_out, T8, dtype=torch.float32, device=dev)
            act = torch.empty(B, self.enc_size, T8, dtype=torch.float32, device=dev)
            ids = torch.empty(B, T8, dtype=torch.int32, device=dev) if self.enc_mode == 'one_hot' else None
            nbytes = lib.zs_encoder_workspace_bytes(h, B, T)
            ws = self._get_workspace(nbytes, dev)
            _lib.check(lib.zs_encoder_forward(h, _ptr(x), B, T, _ptr(noise), _ptr(logits), _ptr(act), _ptr(ids),
                                              _ptr(ws), ws.numel(), _stream()))
        return act, logits, ids

    def forward(self, x, noise=None):
        act, logits, _ = self.encode(x, noise)
        return act, logits


class Decoder(_Packed):
    def __init__(self, c_in=512, c_out=513, c_h=512, c_a=8, ns=0.2, seg_len=64, output_mask=False):
        super().__init__()
        self.output_mask, self.ns, self.seg_len = output_mask, ns, seg_len
        self.c_in, self.c_out, self.c_h, self.c_a = c_in, c_out, c_h, c_a
        for j in range(1, 7):
            setattr(self, f'conv{j}', nn.Conv1d(c_h, 2 * c_h if j % 2 == 1 else c_h, kernel_size=3))
        for j in range(1, 5):
            setattr(self, f'dense{j}', nn.Linear(c_h, c_h))
        self.RNN = nn.GRU(input_size=c_h, hidden_size=c_h // 2, num_layers=1, bidirectional=True)
        self.dense5 = nn.Linear(2 * c_h + c_h, c_h)
        self.linear = nn.Linear(c_h, c_out)
        self.input_emb = nn.Linear(c_in, c_h)
        for j in range(1, 6):
            setattr(self, f'emb{j}', nn.Embedding(c_a, c_h))

    def _free(self):
        if self._handle is not None:
            _lib.lib().zs_decoder_free(self._handle)
            self._handle = None

    def _pack(self, dev):
        lib = _lib.lib()
        cfg = _lib.DecoderCfg(self.c_in, self.c_out, self.c_h, self.c_a, self.seg_len, int(bool(self.output_mask)),
                              _lib.OPERANDS[self.operand], self.ns)
        for p in self.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('Decoder parameters must be contiguous float32')
        w = _lib.DecoderWeights()
        for i in range(6):
            conv = getattr(self, f'conv{i + 1}')
            w.conv_w[i] = conv.weight.data_ptr()
            w.conv_b[i] = conv.bias.data_ptr()
        for i in range(4):
            d = getattr(self, f'dense{i + 1}')
            w.dense_w[i] = d.weight.data_ptr()
            w.dense_b[i] = d.bias.data_ptr()
        for i, sfx in enumerate(('', '_reverse')):
            w.gru_w_ih[i] = getattr(self.RNN, 'weight_ih_l0' + sfx).data_ptr()
            w.gru_w_hh[i] = getattr(self.RNN, 'weight_hh_l0' + sfx).data_ptr()
            w.gru_b_ih[i] = getattr(self.RNN, 'bias_ih_l0' + sfx).data_ptr()
            w.gru_b_hh[i] = getattr(self.RNN, 'bias_hh_l0' + sfx).data_ptr()
        w.dense5_w, w.dense5_b = self.dense5.weight.data_ptr(), self.dense5.bias.data_ptr()
        w.linear_w, w.linear_b = self.linear.weight.data_ptr(), self.linear.bias.data_ptr()
        w.input_emb_w, w.input_emb_b = self.input_emb.weight.data_ptr(), self.input_emb.bias.data_ptr()
        for i in range(5):
            w.emb[i] = getattr(self, f'emb{i + 1}').weight.data_ptr()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.zs_decoder_pack(C.byref(cfg), C.byref(w), _stream(), C.byref(h)))
        self._handle = h

    @torch.no_grad()
    def decode(self, x=None, c=None, unit_ids=None, out=None, accumulate=0):
        """Decoder.forward with the extras the batched front-end uses:
        `unit_ids` (B, T8) int32 replaces a one-hot `x` (input_emb becomes a gather);
        `out`/`accumulate` fuse the patcher combine rules of trainer.py:206-211
        (1: out += y, 2: out += out * y)."""
        src = x if x is not None else unit_ids
        self._check_input(src, 'x')
        dev = src.device
        if x is not None:
            if x.dim() != 3 or x.shape[1] != self.c_in:
                raise RuntimeError(f'Decoder: expected (B, {self.c_in}, T8), got {tuple(x.shape)}')
            x = x.detach().contiguous().float()
            B, _, T8 = x.shape
        else:
            unit_ids = unit_ids.to(torch.int32).contiguous()
            B, T8 = unit_ids.shape
        if not c.is_cuda and c.numel() and (int(c.min()) < 0 or int(c.max()) >= self.c_a):
            raise RuntimeError(f'Decoder: speaker id outside [0, {self.c_a})')   # host ids are validated here;
        c = c.to(dev, torch.int64).contiguous().view(-1)                        # device ids are clamped by the kernel
        if c.numel() != B:
            raise RuntimeError(f'Decoder: {c.numel()} speaker ids for {B} segments')
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._ensure_packed(dev)
            if out is None:
                if accumulate:
                    raise RuntimeError('Decoder: accumulate needs `out`')
                out = torch.empty(B, self.c_out, 8 * T8, dtype=torch.float32, device=dev)
            elif tuple(out.shape) != (B, self.c_out, 8 * T8) or out.dtype != torch.float32 or not out.is_contiguous():
                raise RuntimeError('Decoder: `out` must be a contiguous float32 (B, c_out, 8*T8) tensor')
            nbytes = lib.zs_decoder_workspace_bytes(h, B, T8)
            ws = self._get_workspace(nbytes, dev)
            _lib.check(lib.zs_decoder_forward(h, _ptr(x), _ptr(unit_ids) if x is None else C.c_void_p(0), _ptr(c),
                                              B, T8, _ptr(out), accumulate, _ptr(ws), ws.numel(), _stream()))
        return out

    def forward(self, x, c):
        return self.decode(x, c)
