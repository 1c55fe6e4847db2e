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
  * `forward` is the inference path (no autograd graph, dropout not applied); the pretrain_AE training step
    (trainer.py:321-332) runs through `forward_train` / `backward` below - driven by `zs_b200.train.PretrainAE`
    or, for code that calls `loss.backward()` itself, by the autograd wrappers `zs_b200.train.encode_step /
    decode_step`;
  * segments are limited to 9 <= T <= 256 frames (the range convert.py's chunking produces
    for seg_len = 128); both padding modes of pad_layer (reflect for hps seg_len >= 64, zero below - the latter
    inference only);
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


class OperandRangeError(RuntimeError):
    """fp16 operands saturated (|value| > 65504 clamped in a GEMM epilogue): results are degraded."""


def check_range(device=None, raise_on_saturation=True):
    """Number of fp16 values the GEMM epilogues had to clamp on `device` since the last check (synchronises the current
    stream; resets the counter).  The reference computes in fp32 (model/model.py:20-110) and cannot overflow; this path
    stores activations as fp16 operands, which every InstanceNorm keeps O(1) - a trained checkpoint with a larger
    pre-normalisation range must be run with `module.operand = 'bf16'`.  Never silent: the front-end calls this after every
    batch it brings back to the host."""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        n = _lib.saturation_count(torch.cuda.current_stream().cuda_stream, reset=True)
    if n and raise_on_saturation:
        raise OperandRangeError(f'{n} GEMM epilogue threads clamped fp16 operands to +-65504 on {dev}: the activations of this '
                                "checkpoint exceed the fp16 range - set `encoder.operand = decoder.operand = 'bf16'`")
    return n


class _Packed(nn.Module):
    """Shared plumbing: lazy packing of the fp32 parameters into the library handle."""

    def __init__(self):
        super().__init__()
        self._handle = None
        self._packed_key = None
        self._workspace = None
        self.operand = 'fp16'
        # training handle (cfg.train = 1): re-packed in place after every optimiser step
        self._thandle = None
        self._tpacked_key = None
        self._tworkspace = None
        self._train_ctx = None
        self._train_ctx_version = 0     # bumped by every forward_train: a backward of an older forward is refused

    def _param_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters()) + (self.operand,)

    def _free(self):
        raise NotImplementedError

    def _pack(self, dev):
        raise NotImplementedError

    def _ensure_packed(self, dev):
        key = self._param_key()
        if self._handle is None or key != self._packed_key:
            self._free_eval()
            self._handle = self._pack(dev, train=False)
            self._packed_key = key
        return self._handle

    def _ensure_train_packed(self, dev):
        if self.operand != 'fp16':
            raise RuntimeError('the training path computes in fp16 operands')
        key = self._param_key()
        if self._thandle is None:
            self._thandle = self._pack(dev, train=True)
        elif key != self._tpacked_key:
            self._repack(dev)
        self._tpacked_key = key
        return self._thandle

    def mark_repacked(self):
        self._tpacked_key = self._param_key()

    def _get_train_workspace(self, nbytes, dev):
        ws = self._tworkspace
        if ws is None or ws.numel() < nbytes or ws.device != dev:
            self._tworkspace = ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return ws

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
        elif enc_mode == 'binary':         # model/model.py:391-392: an enc_size x enc_size projection per frame
            if enc_size > 128:
                raise NotImplementedError("enc_mode 'binary' projects to enc_size^2 channels: supported up to enc_size 128")
            n_out = enc_size * enc_size
        else:
            raise NotImplementedError('Invalid encoding mode!')
        self.n_out = n_out
        self.linear = nn.Linear(c_h2 + 2 * c_h3, n_out)

    # -- library handle ------------------------------------------------------------------
    def _free_eval(self):
        if self._handle is not None:
            _lib.lib().zs_encoder_free(self._handle)
            self._handle = None

    def _free(self):
        self._free_eval()
        if self._thandle is not None:
            _lib.lib().zs_encoder_free(self._thandle)
            self._thandle = None

    def weight_table(self, tensors=None):
        """`zs_encoder_weights` filled with the data pointers of the parameters, or of same-named tensors in
        `tensors` (a dict name -> tensor shaped like the parameter: the gradient table of the backward pass)."""
        get = (dict(self.named_parameters()) if tensors is None else tensors).__getitem__
        w = _lib.EncoderWeights()
        for i in range(7):
            w.conv1s_w[i] = get(f'conv1s.{i}.weight').data_ptr()
            w.conv1s_b[i] = get(f'conv1s.{i}.bias').data_ptr()
            w.conv_w[i] = get(f'conv{i + 2}.weight').data_ptr()
            w.conv_b[i] = get(f'conv{i + 2}.bias').data_ptr()
        for i in range(4):
            w.dense_w[i] = get(f'dense{i + 1}.weight').data_ptr()
            w.dense_b[i] = get(f'dense{i + 1}.bias').data_ptr()
        for i, sfx in enumerate(('', '_reverse')):
            w.gru_w_ih[i] = get('RNN.weight_ih_l0' + sfx).data_ptr()
            w.gru_w_hh[i] = get('RNN.weight_hh_l0' + sfx).data_ptr()
            w.gru_b_ih[i] = get('RNN.bias_ih_l0' + sfx).data_ptr()
            w.gru_b_hh[i] = get('RNN.bias_hh_l0' + sfx).data_ptr()
        w.linear_w = get('linear.weight').data_ptr()
        w.linear_b = get('linear.bias').data_ptr()
        return w

    def _pack(self, dev, train=False):
        lib = _lib.lib()
        cfg = _lib.EncoderCfg(self.c_in, self.c_h1, self.c_h2, self.c_h3, self.enc_size,
                              _lib.ENC_MODES[self.enc_mode], self.seg_len, _lib.OPERANDS[self.operand], self.ns,
                              int(train))
        for p in self.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('Encoder parameters must be contiguous float32')
        w = self.weight_table()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.zs_encoder_pack(C.byref(cfg), C.byref(w), _stream(), C.byref(h)))
        return h

    def _repack(self, dev):
        w = self.weight_table()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().zs_encoder_repack(self._thandle, C.byref(w), _stream()))

    # -- training path (model/model.py:440-489 in train mode; trainer.py:246-249 encode_step) ----------
    def forward_train(self, x, noise, dropout_seed=0, keep_masks=None, seed_dev=None):
        """Train-mode forward.  `noise` = Gumbel noise (B, T8, enc_size) on the device; `keep_masks` = optional
        list of six (B, c_h2, T_l) uint8 tensors replaying explicit Dropout draws.  Returns (act, logits, ids);
        the activations the backward needs stay in the module's training workspace until `backward`."""
        self._check_input(x, 'x')
        B, _, T = x.shape
        dev = x.device
        x = x.detach().contiguous().float()
        noise = noise.to(dev, torch.float32).contiguous()
        if tuple(noise.shape) != (B, self.t8(T), self.enc_size):
            raise RuntimeError(f'Encoder.forward_train: noise must be (B, T8, enc_size), got {tuple(noise.shape)}')
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._ensure_train_packed(dev)
            T8 = self.t8(T)
            logits = torch.empty(B, self.n_out, T8, dtype=torch.float32, device=dev)
            act = torch.empty(B, self.enc_size, T8, dtype=torch.float32, device=dev)
            ids = torch.empty(B, T8, dtype=torch.int32, device=dev)
            ws = self._get_train_workspace(lib.zs_encoder_train_workspace_bytes(h, B, T), dev)
            km = self._mask_table(keep_masks)
            _lib.check(lib.zs_encoder_forward_train(h, _ptr(x), B, T, _ptr(noise), float(self.dp), int(dropout_seed) & (2 ** 64 - 1),
                                                    _ptr(seed_dev), km, _ptr(logits), _ptr(act), _ptr(ids), _ptr(ws), ws.numel(),
                                                    _stream()))
        self._train_ctx = (B, T, noise, logits, int(dropout_seed) & (2 ** 64 - 1), keep_masks, seed_dev)
        self._train_ctx_version += 1
        return act, logits, ids

    @staticmethod
    def _mask_table(keep_masks):
        if keep_masks is None:
            return None
        if len(keep_masks) != 6:
            raise RuntimeError('keep_masks must hold the six Dropout keep-masks')
        arr = (C.c_void_p * 6)()
        for i, m in enumerate(keep_masks):
            if m.dtype != torch.uint8 or not m.is_cuda or not m.is_contiguous():
                raise RuntimeError('keep_masks must be contiguous CUDA uint8 tensors')
            arr[i] = m.data_ptr()
        return arr

    def backward(self, d_act, grads, loss_scale, d_act_scale=1.0):
        """Backward of the last `forward_train`: `d_act` (B, enc_size, T8) fp32 = dLoss/d(out_act) * d_act_scale;
        parameter gradients are written into the tensors of `grads` (dict name -> fp32 tensor), which must be ZERO on entry."""
        B, T, noise, logits, seed, keep_masks, seed_dev = self._train_ctx
        dev = d_act.device
        d_act = d_act.contiguous().float()
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._thandle
            ws = self._tworkspace
            g = self.weight_table(grads)
            km = self._mask_table(keep_masks)
            _lib.check(lib.zs_encoder_backward(h, _ptr(d_act), float(d_act_scale), _ptr(noise), _ptr(logits), B, T,
                                               float(self.dp), seed, _ptr(seed_dev), km, float(loss_scale), C.byref(g),
                                               _ptr(ws), ws.numel(), _stream()))

    @staticmethod
    def t8(T):
        return (((T + 1) // 2 + 1) // 2 + 1) // 2

    def noise_shape(self, B, T):
        T8 = self.t8(T)
        return {'one_hot': (B, T8, self.enc_size), 'multilabel_binary': (B, T8, self.enc_size, 2),
                'gumbel_t': (B, self.enc_size, T8), 'binary': (B, T8, self.enc_size, self.enc_size),
                'continues': None}[self.enc_mode]

    @torch.no_grad()
    def encode(self, x, noise=None, layout='nct', noise_seeds=None, want_act=True, want_logits=True):
        """forward() plus the unit ids: returns (out_act, out, unit_ids or None).

        Extras of the batched front-end (all optional, defaults = the reference's forward contract):
          x            float32 or float16 (an fp16 upload is bit-identical: the path rounds to fp16 operands first thing);
          layout       'nct' (B, c_in, T) as Encoder.forward gets it, or 'ntc' (B, T, c_in) as Trainer.test_step gets it
                       before its permute (trainer.py:196) - no transpose copy;
          noise_seeds  one_hot only, with noise=None: (B,) int64 device tensor - the Gumbel noise is drawn ON THE DEVICE, one
                       counter-based stream per segment (same distribution, not the reference's CPU-generator stream;
                       saves 4 KB of upload per unit frame; a segment's draw does not depend on its batch);
          want_act / want_logits = False skip those outputs (one_hot: the ids say everything)."""
        self._check_input(x, 'x')
        if layout not in ('nct', 'ntc'):
            raise RuntimeError(f"Encoder: layout must be 'nct' or 'ntc', got {layout!r}")
        c_axis = 1 if layout == 'nct' else 2
        if x.dim() != 3 or x.shape[c_axis] != self.c_in:
            want = f'(B, {self.c_in}, T)' if layout == 'nct' else f'(B, T, {self.c_in})'
            raise RuntimeError(f'Encoder: expected {want}, got {tuple(x.shape)}')
        B, T = x.shape[0], x.shape[3 - c_axis]
        dev = x.device
        x = x.detach()
        if x.dtype not in (torch.float32, torch.float16):
            x = x.float()
        if x.dtype == torch.float16 and self.operand != 'fp16':
            x = x.float()                                # bf16 operands: round once, from fp32
        x = x.contiguous()
        shape = self.noise_shape(B, T)
        device_noise = False
        if shape is not None:
            if noise is None and noise_seeds is not None:
                if self.enc_mode != 'one_hot':
                    raise RuntimeError('Encoder: device-generated Gumbel noise exists for enc_mode one_hot only')
                if noise_seeds.dtype != torch.int64 or noise_seeds.numel() != B or not noise_seeds.is_cuda:
                    raise RuntimeError('Encoder: noise_seeds must be a (B,) int64 CUDA tensor')
                noise_seeds = noise_seeds.contiguous()
                device_noise = True
            else:
                if noise is None:
                    noise = sample_gumbel(shape)            # CPU generator, like the reference
                if tuple(noise.shape) != shape:
                    raise RuntimeError(f'Encoder: noise must have shape {shape}, got {tuple(noise.shape)}')
                noise = noise.to(dev, torch.float32, non_blocking=True).contiguous()
        else:
            noise = None
        ids_only = self.enc_mode == 'one_hot'
        if not ids_only and not (want_act and want_logits):
            raise RuntimeError('Encoder: want_act / want_logits = False need enc_mode one_hot (the ids carry the result)')
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._ensure_packed(dev)
            T8 = self.t8(T)
            logits = torch.empty(B, self.n_out, T8, dtype=torch.float32, device=dev) if want_logits else None
            act = torch.empty(B, self.enc_size, T8, dtype=torch.float32, device=dev) if want_act else None
            ids = torch.empty(B, T8, dtype=torch.int32, device=dev) if self.enc_mode == 'one_hot' else None
            nbytes = lib.zs_encoder_workspace_bytes(h, B, T)
            ws = self._get_workspace(nbytes, dev)
            _lib.check(lib.zs_encoder_forward_x(h, _ptr(x), 1 if x.dtype == torch.float16 else 0, 0 if layout == 'nct' else 1, B, T,
                                                _ptr(None if device_noise else noise), _ptr(noise_seeds if device_noise else None),
                                                _ptr(logits), _ptr(act), _ptr(ids), _ptr(ws), ws.numel(), _stream()))
        return act, logits, ids

    def forward(self, x, noise=None):
        """model/model.py:440-489.  eval() / no_grad: the inference kernels.  train() with autograd enabled (the
        reference's `encode_step`, trainer.py:246-249): the training forward behind a `torch.autograd.Function`, so
        the reference's own loop (loss.backward(), grad_clip, ae_opt.step()) runs unchanged on these modules."""
        if self.training and torch.is_grad_enabled():
            from .train import encode_step
            return encode_step(self, x, noise)
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

    def _free_eval(self):
        if self._handle is not None:
            _lib.lib().zs_decoder_free(self._handle)
            self._handle = None

    def _free(self):
        self._free_eval()
        if self._thandle is not None:
            _lib.lib().zs_decoder_free(self._thandle)
            self._thandle = None

    def weight_table(self, tensors=None):
        """`zs_decoder_weights` of the parameters, or of same-named tensors in `tensors` (gradient table)."""
        get = (dict(self.named_parameters()) if tensors is None else tensors).__getitem__
        w = _lib.DecoderWeights()
        for i in range(6):
            w.conv_w[i] = get(f'conv{i + 1}.weight').data_ptr()
            w.conv_b[i] = get(f'conv{i + 1}.bias').data_ptr()
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
        w.input_emb_w, w.input_emb_b = get('input_emb.weight').data_ptr(), get('input_emb.bias').data_ptr()
        for i in range(5):
            w.emb[i] = get(f'emb{i + 1}.weight').data_ptr()
        return w

    def _pack(self, dev, train=False):
        lib = _lib.lib()
        cfg = _lib.DecoderCfg(self.c_in, self.c_out, self.c_h, self.c_a, self.seg_len, int(bool(self.output_mask)),
                              _lib.OPERANDS[self.operand], self.ns, int(train))
        for p in self.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('Decoder parameters must be contiguous float32')
        w = self.weight_table()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.zs_decoder_pack(C.byref(cfg), C.byref(w), _stream(), C.byref(h)))
        return h

    def _repack(self, dev):
        w = self.weight_table()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().zs_decoder_repack(self._thandle, C.byref(w), _stream()))

    # -- training path (model/model.py:344-365; trainer.py:251-254 decode_step) -------------------------
    def forward_train(self, x, c):
        """Train-mode forward on dense activations x (B, c_in, T8); keeps what `backward` needs."""
        self._check_input(x, 'x')
        dev = x.device
        x = x.detach().contiguous().float()
        B, _, T8 = x.shape
        c = c.to(dev, torch.int64).contiguous().view(-1)
        if c.numel() != B:
            raise RuntimeError(f'Decoder: {c.numel()} speaker ids for {B} segments')
        lib = _lib.lib()
        with torch.cuda.device(dev):
            h = self._ensure_train_packed(dev)
            spec = torch.empty(B, self.c_out, 8 * T8, dtype=torch.float32, device=dev)
            ws = self._get_train_workspace(lib.zs_decoder_train_workspace_bytes(h, B, T8), dev)
            _lib.check(lib.zs_decoder_forward_train(h, _ptr(x), _ptr(c), B, T8, _ptr(spec), _ptr(ws), ws.numel(), _stream()))
        self._train_ctx = (B, T8, c, spec)
        self._train_ctx_version += 1
        return spec

    def backward(self, grads, loss_scale, target=None, d_spec=None, loss_out=None, want_d_act=True):
        """Backward of the last `forward_train`.  Either `target` (B, c_out, T): the L1 loss of trainer.py:327 is fused
        and added to `loss_out` (device fp32 scalar, zero it first), or `d_spec` = dLoss/dspec.  Parameter gradients
        are written into `grads` (zero on entry); returns d_act = dLoss/d(enc_act) * loss_scale, (B, c_in, T8) fp32."""
        B, T8, c, spec = self._train_ctx
        dev = spec.device
        lib = _lib.lib()
        if target is not None:
            target = target.contiguous().float()
            if tuple(target.shape) != tuple(spec.shape):
                raise RuntimeError(f'Decoder.backward: target {tuple(target.shape)} vs output {tuple(spec.shape)}')
        if d_spec is not None:
            d_spec = d_spec.contiguous().float()
        with torch.cuda.device(dev):
            d_act = torch.empty(B, self.c_in, T8, dtype=torch.float32, device=dev) if want_d_act else None
            g = self.weight_table(grads)
            ws = self._tworkspace
            _lib.check(lib.zs_decoder_backward(self._thandle, _ptr(spec), _ptr(target), _ptr(d_spec), _ptr(c), B, T8,
                                               float(loss_scale), _ptr(loss_out), C.byref(g), _ptr(d_act), _ptr(ws),
                                               ws.numel(), _stream()))
        return d_act

    @torch.no_grad()
    def decode(self, x=None, c=None, unit_ids=None, out=None, accumulate=0, out_dtype=torch.float32):
        """Decoder.forward with the extras the batched front-end uses:
        `unit_ids` (B, T8) int32 replaces a one-hot `x` (input_emb becomes a gather);
        `out`/`accumulate` fuse the patcher combine rules of trainer.py:206-211
        (1: out += y, 2: out += out * y);
        `out_dtype` float16 rounds the (0, 1) output once to fp16 (<= 2.5e-4 absolute): half the download bytes."""
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
                if out_dtype not in (torch.float32, torch.float16):
                    raise RuntimeError('Decoder: out_dtype must be float32 or float16')
                out = torch.empty(B, self.c_out, 8 * T8, dtype=out_dtype, device=dev)
            elif tuple(out.shape) != (B, self.c_out, 8 * T8) or out.dtype not in (torch.float32, torch.float16) or not out.is_contiguous():
                raise RuntimeError('Decoder: `out` must be a contiguous float32 / float16 (B, c_out, 8*T8) tensor')
            nbytes = lib.zs_decoder_workspace_bytes(h, B, T8)
            ws = self._get_workspace(nbytes, dev)
            _lib.check(lib.zs_decoder_forward_x(h, _ptr(x), _ptr(unit_ids) if x is None else C.c_void_p(0), _ptr(c),
                                                B, T8, _ptr(out), 1 if out.dtype == torch.float16 else 0, accumulate, _ptr(ws),
                                                ws.numel(), _stream()))
        return out

    def forward(self, x, c):
        """model/model.py:344-365; train() with autograd enabled = `decode_step` (trainer.py:251-254) with a graph."""
        if self.training and torch.is_grad_enabled():
            from .train import decode_step
            return decode_step(self, x, c)
        return self.decode(x, c)
