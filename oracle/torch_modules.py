"""ORACLE - test / baseline infrastructure, NOT product code.

`nn.Module` restatement of the reference's Encoder / Decoder (model/model.py:283-489) built from stock torch layers
(nn.Conv1d, nn.Linear, nn.GRU, F.instance_norm ...) with the reference's parameter names, so a reference-layout
``state_dict`` loads with ``strict=True``.  It exists for ONE purpose: `bench.py`'s ``cuda_eager_baseline`` - the
"existing Blackwell path" SURVEY.md 8(d) asks to be timed next to ours, i.e. what a user of the reference gets by
calling ``.cuda()`` on its modules: cuDNN convolutions (TF32 by torch's default), cuBLAS linears, the cuDNN GRU, one
ATen launch per elementwise op.  `oracle/ae_oracle.py` (plain functions, explicit GRU loop) stays the parity oracle;
`tests/test_oracle_golden.py` holds these modules to it on the CPU.

The Gumbel noise is an explicit argument (the reference draws it on the CPU inside forward, model/model.py:96-98, and
uploads it - the baseline does the same upload when handed a CPU tensor).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

GUMBEL_TAU = 0.1


def _same_pad(x, k, mode):
    # model/model.py:20-40: (k//2, k//2) for odd k, (k//2, k//2 - 1) for even k
    lo, hi = (k // 2, k // 2) if k % 2 else (k // 2, k // 2 - 1)
    return F.pad(x, (lo, hi), mode=mode) if k > 1 else x


def _per_frame(layer, x):
    # model/model.py:69-78 linear(): nn.Linear applied along the channel axis of (B, C, T)
    return layer(x.transpose(1, 2)).transpose(1, 2)


def _bi_gru(rnn, x):
    # model/model.py:59-66 RNN(): (B, C, T) -> (T, B, C), zero initial state, back to (B, 2H, T)
    rnn.flatten_parameters()
    out, _ = rnn(x.permute(2, 0, 1))
    return out.permute(1, 2, 0)


class TorchEncoder(nn.Module):
    """model/model.py:368-489, eval-mode forward, enc_mode one_hot / continues."""

    def __init__(self, c_in=513, c_h1=128, c_h2=512, c_h3=128, ns=0.01, enc_size=1024, seg_len=128, enc_mode='one_hot'):
        super().__init__()
        self.ns, self.enc_mode = ns, enc_mode
        self.pad_mode = 'reflect' if seg_len >= 64 else 'constant'       # model/model.py:38
        self.conv1s = nn.ModuleList(nn.Conv1d(c_in, c_h1, k) for k in range(1, 8))
        self.conv2 = nn.Conv1d(7 * c_h1 + c_in, c_h2, 1)
        for j in range(3, 9):
            setattr(self, f'conv{j}', nn.Conv1d(c_h2, c_h2, 5, stride=1 + (j % 2 == 0)))
        for j in range(1, 5):
            setattr(self, f'dense{j}', nn.Linear(c_h2, c_h2))
        self.RNN = nn.GRU(c_h2, c_h3, bidirectional=True)
        self.linear = nn.Linear(c_h2 + 2 * c_h3, enc_size)

    def _conv(self, name, x):
        layer = getattr(self, name)
        return F.leaky_relu(layer(_same_pad(x, layer.kernel_size[0], self.pad_mode)), self.ns)

    def forward(self, x, noise=None):
        bank = [c(_same_pad(x, c.kernel_size[0], self.pad_mode)) for c in self.conv1s]          # :441-446
        h = F.leaky_relu(torch.cat(bank + [x], 1), self.ns)
        h = F.instance_norm(self._conv('conv2', h))                                             # :447
        for a, b in (('conv3', 'conv4'), ('conv5', 'conv6'), ('conv7', 'conv8')):                # :448-450
            y = F.instance_norm(self._conv(b, self._conv(a, h)))
            res = F.pad(h, (0, h.shape[2] % 2), mode=self.pad_mode)
            h = y + F.avg_pool1d(res, 2)
        for a, b in (('dense1', 'dense2'), ('dense3', 'dense4')):                                # :452-453
            y = F.leaky_relu(_per_frame(getattr(self, a), h), self.ns)
            y = F.leaky_relu(_per_frame(getattr(self, b), y), self.ns)
            h = F.instance_norm(y) + h
        h = torch.cat([h, _bi_gru(self.RNN, h)], 1)                                              # :454-455
        logits = _per_frame(self.linear, h)
        if self.enc_mode == 'continues':
            return F.leaky_relu(logits, self.ns), logits
        # :461-464 + :93-110: softmax((l + g) / tau), hard one-hot of its argmax
        y = F.softmax((logits.transpose(1, 2) + noise.to(logits.device)) / GUMBEL_TAU, dim=-1)
        hard = torch.zeros_like(y).scatter_(-1, y.argmax(-1, keepdim=True), 1.0)
        return hard.transpose(1, 2).contiguous(), logits


class TorchDecoder(nn.Module):
    """model/model.py:283-365."""

    def __init__(self, c_in=1024, c_out=513, c_h=1024, c_a=102, ns=0.01, seg_len=128, output_mask=False):
        super().__init__()
        self.ns, self.output_mask = ns, output_mask
        self.pad_mode = 'reflect' if seg_len >= 64 else 'constant'
        for j in range(1, 7):
            setattr(self, f'conv{j}', nn.Conv1d(c_h, c_h * (1 + j % 2), 3))
        for j in range(1, 5):
            setattr(self, f'dense{j}', nn.Linear(c_h, c_h))
        self.RNN = nn.GRU(c_h, c_h // 2, bidirectional=True)
        self.dense5 = nn.Linear(3 * c_h, c_h)
        self.linear = nn.Linear(c_h, c_out)
        self.input_emb = nn.Linear(c_in, c_h)
        for j in range(1, 6):
            setattr(self, f'emb{j}', nn.Embedding(c_a, c_h))

    def forward(self, x, c):
        e = [getattr(self, f'emb{j}')(c).unsqueeze(2) for j in range(1, 6)]
        h = _per_frame(self.input_emb, x)                                                       # :346
        for blk in range(3):                                                                    # :317-331
            up, same, eb = getattr(self, f'conv{2 * blk + 1}'), getattr(self, f'conv{2 * blk + 2}'), e[blk]
            y = F.leaky_relu(up(_same_pad(h + eb, 3, self.pad_mode)), self.ns)
            b, c2, w = y.shape
            y = y.view(b, c2 // 2, 2, w).transpose(2, 3).reshape(b, c2 // 2, 2 * w) + eb         # pixel shuffle :43-51
            y = F.leaky_relu(same(_same_pad(y, 3, self.pad_mode)), self.ns)
            h = F.instance_norm(y) + F.interpolate(h, scale_factor=2, mode='nearest')
        for a, b in (('dense1', 'dense2'), ('dense3', 'dense4')):                                # :333-342, emb4 twice
            y = F.leaky_relu(_per_frame(getattr(self, a), h + e[3]), self.ns)
            y = F.leaky_relu(_per_frame(getattr(self, b), y + e[3]), self.ns)
            h = F.instance_norm(y) + h
        rnn = _bi_gru(self.RNN, h + e[4])                                                        # :352-355
        h = torch.cat([h, rnn, e[4].expand(-1, -1, h.shape[2])], 1)                              # :356-357
        h = F.leaky_relu(_per_frame(self.dense5, h), self.ns)
        h = _per_frame(self.linear, h)
        return torch.tanh(h) if self.output_mask else torch.sigmoid(h)
