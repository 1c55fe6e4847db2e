"""Drop-in `PatchDiscriminator` / `TargetClassifier` (reference model/model.py:113-226), forward passes on libzsae.so.

SURVEY.md section 8 f4, forward part: the stage-2 critic (WGAN value + auxiliary speaker logits, `Trainer.patch_step`,
trainer.py:257-265) and the target-speaker classifier behind `Trainer.classify` (trainer.py:230-235).  Same constructor
signatures, parameter names and shapes as the reference (`load_state_dict(strict=True)` of a reference checkpoint works).

Every 5x5 stride-2 Conv2d runs on the tcgen05 conv GEMM as a 1-D stride-2 convolution along time whose channels are the
five kernel rows (include/zs_ae.h: zs_conv2d_gather -> zs_conv1d_cl -> zs_instnorm2d_stats); the InstanceNorm2d of layer
i is applied by the gather of layer i + 1.  Inference only: Dropout2d masks, the backward pass and the WGAN-GP double
backward of the adversarial step (trainer.py:467-560, utils.py:58-77) are not built - a train-mode forward raises.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .model import _ptr, _stream

_KW = {128: 4, 64: 2, 32: 1}          # time extent of conv7 / conv_classify (model/model.py:123-131)
_CHANS = [(1, 64), (64, 128), (128, 256), (256, 512), (512, 512)]


def _round_up(v, m):
    return (v + m - 1) // m * m


class _Critic2d(nn.Module):
    """The six conv blocks both networks share (model/model.py:146-160, 205-218) + the full-map heads."""

    def __init__(self, n_class, ns, dp, seg_len):
        super().__init__()
        if seg_len not in _KW:
            raise NotImplementedError('Segement length {} is not supported!'.format(seg_len))      # model/model.py:132
        self.ns, self.dp, self.seg_len, self.n_class = ns, dp, seg_len, n_class
        for i, (ci, co) in enumerate(_CHANS, 1):
            setattr(self, f'conv{i}', nn.Conv2d(ci, co, kernel_size=5, stride=2))
        self.conv6 = nn.Conv2d(512, 32, kernel_size=1)
        self.conv7 = nn.Conv2d(32, 1, kernel_size=(17, _KW[seg_len]))
        self.conv_classify = nn.Conv2d(32, n_class, kernel_size=(17, _KW[seg_len]))
        self._packed = None
        self._packed_key = None
        self._bufs = {}

    # ---- packing: fp32 parameters -> fp16 GEMM operands (once per load_state_dict / device) -----------------------
    def _param_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def _ensure_packed(self, dev):
        key = self._param_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        layers = []
        for i in range(1, 7):
            conv = getattr(self, f'conv{i}')
            w = conv.weight.detach().to(dev, torch.float32)               # (co, ci, kh, kw)
            co, ci, kh, kw = w.shape
            m_rows, c_in = _round_up(co, 128), kh * ci
            c_pad = _round_up(c_in, 64)
            wp = torch.zeros(m_rows, kw, c_pad, dtype=torch.float16, device=dev)
            wp[:co, :, :c_in] = w.permute(0, 3, 2, 1).reshape(co, kw, c_in).to(torch.float16)      # [co][tap kw][kh * ci + c]
            bias = torch.zeros(m_rows, dtype=torch.float32, device=dev)
            bias[:co] = conv.bias.detach().to(dev, torch.float32)
            layers.append(dict(w=wp.contiguous(), bias=bias, m_rows=m_rows, c_out=co, c_in=ci, kh=kh, taps=kw, c_pad=c_pad))
        heads = {}
        for name, conv in (('value', self.conv7), ('classify', self.conv_classify)):
            w = conv.weight.detach().to(dev, torch.float32)               # (J, 32, 17, kw)
            heads[name] = (w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous(),             # [J][p = h * kw + w][c]
                           conv.bias.detach().to(dev, torch.float32).contiguous())
        w_all = torch.cat([heads['value'][0], heads['classify'][0]]).contiguous()
        b_all = torch.cat([heads['value'][1], heads['classify'][1]]).contiguous()
        self._packed = dict(layers=layers, head_w=w_all, head_b=b_all)
        self._packed_key = key
        return self._packed

    def _work_buffers(self, B, T, dev):
        key = (B, T, str(dev))
        if key in self._bufs:
            return self._bufs[key]
        H, W, geo = 513, T, []
        for ci, co in _CHANS:
            Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            geo.append(dict(H=H, W=W, Ho=Ho, Wo=Wo, Wp=W + 4,
                            G=torch.empty(B * Ho, W + 4, _round_up(5 * ci, 8), dtype=torch.float16, device=dev),
                            Y=torch.empty(B * Ho, Wo, co, dtype=torch.float16, device=dev),
                            stats=torch.empty(B, co, 2, dtype=torch.int64, device=dev)))
            H, W = Ho, Wo
        bufs = dict(geo=geo, H6=H, W6=W,
                    G6=torch.empty(B, H * W, 512, dtype=torch.float16, device=dev),
                    Y6=torch.empty(B, H * W, 32, dtype=torch.float16, device=dev),
                    out=torch.empty(B, 1 + self.n_class, dtype=torch.float32, device=dev))
        self._bufs = {key: bufs}
        return bufs

    @staticmethod
    def _conv(layer, src, in_rows, in_pitch, c_in_valid, stride, n_seg, T_out, out, ns, inorm):
        d = _lib.ConvDesc()
        d.w, d.m_rows, d.m_valid, d.taps, d.c_in_pad, d.w_taps, d.bank = layer['w'].data_ptr(), layer['m_rows'], layer['c_out'], layer['taps'], layer['c_pad'], layer['taps'], 0
        d.in_, d.in_rows, d.in_pitch, d.in_row0, d.c_in_valid = src.data_ptr(), in_rows, in_pitch, 0, c_in_valid
        d.stride, d.B, d.T_out = stride, n_seg, T_out
        d.bias, d.spk, d.n_spk, d.lrelu, d.ns, d.inorm = layer['bias'].data_ptr(), None, 1, 1, ns, int(inorm)
        d.res_mode, d.res = 0, None
        d.act, d.out_mode = 0, 0
        d.out, d.out_rows, d.out_pitch, d.out_halo, d.out_choff = out.data_ptr(), T_out, layer['c_out'], 0, 0
        d.accumulate, d.operand, d.nb_hint, d.out_f16 = 0, _lib.OPERANDS['fp16'], 0, 0
        _lib.check(_lib.lib().zs_conv1d_cl(C.byref(d), _stream()))

    @torch.no_grad()
    def _heads(self, x):
        """x: (B, 513, T) fp32 cuda -> (B, 1 + n_class) fp32: column 0 = conv7 value, the rest = conv_classify logits."""
        if self.training and self.dp > 0:
            raise NotImplementedError(f'{type(self).__name__}: the train-mode forward (Dropout2d + the adversarial backward, '
                                      'trainer.py:467-560) is not built - call .eval()')
        if self.seg_len < 64:
            raise NotImplementedError(f'{type(self).__name__}: seg_len {self.seg_len} pads with zeros (model/model.py:38); only reflect padding is built')
        if not x.is_cuda:
            raise RuntimeError(f'{type(self).__name__} runs on libzsae.so (CUDA): move the input to a B200 (there is no CPU path)')
        if x.dim() != 3 or x.shape[1] != 513 or x.shape[2] != self.seg_len:
            raise ValueError(f'{type(self).__name__}: expected (B, 513, {self.seg_len}), got {tuple(x.shape)}')
        x = x.contiguous().float()
        B, dev = x.shape[0], x.device
        lib = _lib.lib()
        with torch.cuda.device(dev):
            pk = self._ensure_packed(dev)
            bf = self._work_buffers(B, self.seg_len, dev)
            src, src_f32, stats, c_in = x, 1, None, 1
            for layer, g in zip(pk['layers'][:5], bf['geo']):
                _lib.check(lib.zs_conv2d_gather(_ptr(src), src_f32, B, g['H'], g['W'], c_in, c_in, 5, 2, 2, 2, g['Ho'], g['Wp'],
                                                _ptr(stats), 1.0 / (g['H'] * g['W']), _ptr(g['G']), g['G'].shape[2], _stream()))
                # (layer 1: 5 real channels in a pitch of 8, the gather zero-fills the rest)
                self._conv(layer, g['G'], g['Wp'], g['G'].shape[2], g['G'].shape[2], 2, B * g['Ho'], g['Wo'], g['Y'], self.ns, False)
                _lib.check(lib.zs_instnorm2d_stats(_ptr(g['Y']), B, g['Ho'] * g['Wo'], layer['c_out'], layer['c_out'], _ptr(g['stats']), _stream()))
                src, src_f32, stats, c_in = g['Y'], 0, g['stats'], layer['c_out']
            H6, W6, P6 = bf['H6'], bf['W6'], bf['H6'] * bf['W6']
            # conv6 (1x1) on the normalised map, its InstanceNorm over the P6 positions of a sample fused into the GEMM epilogue
            _lib.check(lib.zs_conv2d_gather(_ptr(src), 0, B, H6, W6, 512, 512, 1, 1, 0, 0, H6, W6, _ptr(stats), 1.0 / P6, _ptr(bf['G6']), 512, _stream()))
            self._conv(pk['layers'][5], bf['G6'], P6, 512, 512, 1, B, P6, bf['Y6'], self.ns, True)
            J = 1 + self.n_class
            _lib.check(lib.zs_critic_head(_ptr(bf['Y6']), B, P6, 32, 32, _ptr(pk['head_w']), _ptr(pk['head_b']), J, _ptr(bf['out']), _stream()))
        return bf['out'].clone()


class PatchDiscriminator(_Critic2d):
    """model/model.py:113-166."""

    def __init__(self, n_class=33, ns=0.2, dp=0.1, seg_len=128):
        super().__init__(n_class, ns, dp, seg_len)

    def forward(self, x, classify=False):
        out = self._heads(x)
        mean_val = out[:, 0].contiguous()          # conv7's map is 1 x 1: the mean over it (model/model.py:157-159) is the value
        return (mean_val, out[:, 1:].contiguous()) if classify else mean_val


class TargetClassifier(_Critic2d):
    """model/model.py:169-223 (declares conv7 like the critic, uses conv_classify only)."""

    def __init__(self, n_class=2, ns=0.2, dp=0.8, seg_len=128):
        super().__init__(n_class, ns, dp, seg_len)

    def forward(self, x):
        return self._heads(x)[:, 1:].contiguous()


def classify(target_classifier, x):
    """`Trainer.classify` (trainer.py:230-235): x (B, T, 513) frames-major features -> logits as a numpy array, eval mode."""
    target_classifier.eval()
    dev = next(target_classifier.parameters()).device
    x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(dev, torch.float32).permute(0, 2, 1)
    return target_classifier(x).cpu().numpy()
