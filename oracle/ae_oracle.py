"""ORACLE - test infrastructure, NOT product code.

CPU fp32 restatement of the reference's autoencoder hot path
(Encoder -> discrete bottleneck -> speaker-conditioned Decoder), written as
plain functions over a reference-layout ``state_dict``.  It exists so the
`-m gpu` parity tests, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline``/``--impl reference`` arm have something to check and time the
CUDA path against on a box where /root/reference does not exist.  Nothing under
``zerospeech-tts-without-t_b200/`` may import it.

Parity pin: the reference has no tests or golden vectors (SURVEY.md section 4,
8c), so this file is pinned against the *live* reference modules
(``/root/reference/model/model.py`` imported in the build container) by
``tests/golden/make_golden.py``, which stores reference outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them against
this restatement (max-abs <= 1e-5, unit ids bit-exact).

Every function cites the reference lines it restates.  All tensors are
(batch, channel, time) float32 like the reference's.
"""
import math

import torch
import torch.nn.functional as F

IN_EPS = 1e-5          # nn.InstanceNorm1d default (model/model.py:303-307, 402-407)
GUMBEL_EPS = 1e-20     # model/model.py:95
GUMBEL_TAU = 0.1       # model/model.py:93


# ----------------------------------------------------------------------------
# helpers: model/model.py:20-110
# ----------------------------------------------------------------------------
def _pad_mode(seg_len):
    # model/model.py:38 - keyed on the hyper-parameter, not on the tensor length
    return 'constant' if seg_len < 64 else 'reflect'


def conv_same(x, w, b, seg_len, stride=1):
    """pad_layer + nn.Conv1d: model/model.py:20-40 (even k pads (k/2, k/2-1))."""
    k = w.shape[2]
    pad = (k // 2, k // 2 - 1) if k % 2 == 0 else (k // 2, k // 2)
    return F.conv1d(F.pad(x, pad, mode=_pad_mode(seg_len)), w, b, stride=stride)


def frame_linear(x, w, b):
    """linear(): per-frame nn.Linear on (B, C, T); model/model.py:69-78."""
    return torch.einsum('oc,bct->bot', w, x) + b.view(1, -1, 1)


def instance_norm(x):
    """nn.InstanceNorm1d, no affine, no running stats, biased variance."""
    mu = x.mean(dim=2, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=2, keepdim=True)
    return (x - mu) / torch.sqrt(var + IN_EPS)


def pixel_shuffle_1d(x):
    """model/model.py:43-51: out[b, c, 2w+r] = in[b, 2c+r, w]."""
    b, c2, w = x.shape
    return x.view(b, c2 // 2, 2, w).permute(0, 1, 3, 2).reshape(b, c2 // 2, 2 * w)


def nearest_up2(x):
    """upsample(): model/model.py:54-56 (always x2)."""
    return x.repeat_interleave(2, dim=2)


def bi_gru(x, sd, prefix='RNN.'):
    """RNN() with zero initial state: model/model.py:59-66, PyTorch gate order r,z,n."""
    B, _, T = x.shape
    outs = []
    for sfx, order in (('', range(T)), ('_reverse', range(T - 1, -1, -1))):
        w_ih, w_hh = sd[f'{prefix}weight_ih_l0{sfx}'], sd[f'{prefix}weight_hh_l0{sfx}']
        b_ih, b_hh = sd[f'{prefix}bias_ih_l0{sfx}'], sd[f'{prefix}bias_hh_l0{sfx}']
        H = w_hh.shape[1]
        gx = torch.einsum('gc,bct->btg', w_ih, x) + b_ih
        h = x.new_zeros(B, H)
        out = x.new_zeros(B, H, T)
        for t in order:
            gh = h @ w_hh.t() + b_hh
            r = torch.sigmoid(gx[:, t, :H] + gh[:, :H])
            z = torch.sigmoid(gx[:, t, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gx[:, t, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            out[:, :, t] = h
        outs.append(out)
    return torch.cat(outs, dim=1)


def gumbel_noise(uniform):
    """_sample_gumbel(): model/model.py:95-98, from an explicit uniform draw."""
    return -torch.log(-torch.log(uniform + GUMBEL_EPS) + GUMBEL_EPS)


def gumbel_hard(logits_last, uniform):
    """gumbel_softmax(): model/model.py:93-110 forward value (exact 0/1) + indices.

    Follows the reference literally: softmax((l+g)/tau) first, then max over the
    last axis (first index wins ties, like torch.max)."""
    y = F.softmax((logits_last + gumbel_noise(uniform)) / GUMBEL_TAU, dim=-1)
    ind = y.max(dim=-1)[1]
    hard = torch.zeros_like(y).scatter_(-1, ind.unsqueeze(-1), 1.0)
    return (hard - y).detach() + y, ind        # :110 straight-through: value = hard, gradient = d softmax


# ----------------------------------------------------------------------------
# Encoder: model/model.py:416-489
# ----------------------------------------------------------------------------
def _drop(out, keep, dp):
    """nn.Dropout in train mode with an explicit keep-mask (None = eval / p = 0)."""
    return out if keep is None else out * keep / (1.0 - dp)


def _enc_conv_block(x, sd, names, strides, ns, seg_len, res=True, keep=None, dp=0.5):
    out = x
    for n, s in zip(names, strides):                         # :418-420
        out = F.leaky_relu(conv_same(out, sd[n + '.weight'], sd[n + '.bias'], seg_len, s), ns)
    out = _drop(instance_norm(out), keep, dp)                 # :421-422 norm_layers = [InstanceNorm, Dropout]
    if res:                                                   # :423-426
        xp = F.pad(x, (0, x.shape[2] % 2), mode=_pad_mode(seg_len))
        out = F.avg_pool1d(xp, 2) + out
    return out


def _enc_dense_block(x, sd, names, ns, keep=None, dp=0.5):
    out = x
    for n in names:                                           # :431-433
        out = F.leaky_relu(frame_linear(out, sd[n + '.weight'], sd[n + '.bias']), ns)
    return _drop(instance_norm(out), keep, dp) + x            # :434-437


def encoder_trunk(sd, x, ns=0.01, seg_len=128, keep_masks=None, dp=0.5):
    """Everything up to (and including) the final Linear: model/model.py:440-455 + linear.

    `keep_masks`: the six Dropout keep-masks (drop1..drop6, 0/1 tensors shaped like the block outputs) of a
    train-mode call; None = eval mode."""
    km = keep_masks if keep_masks is not None else [None] * 6
    bank = [conv_same(x, sd[f'conv1s.{i}.weight'], sd[f'conv1s.{i}.bias'], seg_len) for i in range(7)]
    out = F.leaky_relu(torch.cat(bank + [x], dim=1), ns)      # :445-446
    out = _enc_conv_block(out, sd, ['conv2'], [1], ns, seg_len, res=False, keep=km[0], dp=dp)
    out = _enc_conv_block(out, sd, ['conv3', 'conv4'], [1, 2], ns, seg_len, keep=km[1], dp=dp)
    out = _enc_conv_block(out, sd, ['conv5', 'conv6'], [1, 2], ns, seg_len, keep=km[2], dp=dp)
    out = _enc_conv_block(out, sd, ['conv7', 'conv8'], [1, 2], ns, seg_len, keep=km[3], dp=dp)
    out = _enc_dense_block(out, sd, ['dense1', 'dense2'], ns, keep=km[4], dp=dp)
    out = _enc_dense_block(out, sd, ['dense3', 'dense4'], ns, keep=km[5], dp=dp)
    out = torch.cat([out, bi_gru(out, sd)], dim=1)            # :454-455
    return frame_linear(out, sd['linear.weight'], sd['linear.bias'])


def encoder_forward(sd, x, uniform=None, ns=0.01, seg_len=128, enc_mode='one_hot', enc_size=None, keep_masks=None,
                    dp=0.5):
    """Encoder.forward in eval mode: returns (out_act, out, unit_ids or None).

    `uniform` is the torch.rand draw of gumbel_softmax (shape (B,T8,enc) for one_hot,
    (B,T8,enc,2) for multilabel_binary, (B,T8,enc,enc) for binary, (B,enc,T8) for gumbel_t)."""
    logits = encoder_trunk(sd, x, ns, seg_len, keep_masks, dp)
    ids = None
    if enc_mode == 'continues':                               # :457-459
        act = F.leaky_relu(logits, ns)
    elif enc_mode == 'one_hot':                               # :461-464
        hard, ids = gumbel_hard(logits.permute(0, 2, 1), uniform)
        act = hard.permute(0, 2, 1).contiguous()
    elif enc_mode == 'multilabel_binary':                     # :474-480
        B, C2, T8 = logits.shape
        proj = logits.permute(0, 2, 1).reshape(B, T8, C2 // 2, 2)
        hard, _ = gumbel_hard(proj, uniform)
        act = hard[..., 0].permute(0, 2, 1).contiguous()
    elif enc_mode == 'binary':                                # :466-472
        B, C2, T8 = logits.shape
        E = int(round(C2 ** 0.5))
        proj = logits.permute(0, 2, 1).reshape(B, T8, E, E)
        hard, _ = gumbel_hard(proj, uniform)
        act = torch.clamp(hard.sum(2), min=0, max=1).permute(0, 2, 1).contiguous()
    elif enc_mode == 'gumbel_t':                              # :482-484 (softmax over time)
        act, _ = gumbel_hard(logits, uniform)
    else:
        raise NotImplementedError(enc_mode)
    return act, logits, ids


# ----------------------------------------------------------------------------
# Decoder: model/model.py:317-365
# ----------------------------------------------------------------------------
def _dec_conv_block(x, sd, n1, n2, emb, ns, seg_len):
    e = emb.unsqueeze(2)
    out = F.leaky_relu(conv_same(x + e, sd[n1 + '.weight'], sd[n1 + '.bias'], seg_len), ns)   # :319-321
    out = pixel_shuffle_1d(out) + e                                                           # :323-324
    out = F.leaky_relu(conv_same(out, sd[n2 + '.weight'], sd[n2 + '.bias'], seg_len), ns)     # :325-326
    return instance_norm(out) + nearest_up2(x)                                                # :327-330


def _dec_dense_block(x, sd, names, emb, ns):
    e = emb.unsqueeze(2)
    out = x
    for n in names:                                                                           # :335-338
        out = F.leaky_relu(frame_linear(out + e, sd[n + '.weight'], sd[n + '.bias']), ns)
    return instance_norm(out) + x                                                             # :339-341


def decoder_forward(sd, enc_act, c, ns=0.01, seg_len=128, output_mask=False):
    """Decoder.forward: model/model.py:344-365 (emb4 is used by both dense blocks)."""
    e = [sd[f'emb{j}.weight'][c] for j in range(1, 6)]
    out = frame_linear(enc_act, sd['input_emb.weight'], sd['input_emb.bias'])
    out = _dec_conv_block(out, sd, 'conv1', 'conv2', e[0], ns, seg_len)
    out = _dec_conv_block(out, sd, 'conv3', 'conv4', e[1], ns, seg_len)
    out = _dec_conv_block(out, sd, 'conv5', 'conv6', e[2], ns, seg_len)
    out = _dec_dense_block(out, sd, ['dense1', 'dense2'], e[3], ns)
    out = _dec_dense_block(out, sd, ['dense3', 'dense4'], e[3], ns)
    rnn = bi_gru(out + e[4].unsqueeze(2), sd)                                                 # :352-355
    out = torch.cat([out, rnn, e[4].unsqueeze(2).expand(-1, -1, out.shape[2])], dim=1)        # :356-357
    out = F.leaky_relu(frame_linear(out, sd['dense5.weight'], sd['dense5.bias']), ns)         # :358-359
    out = frame_linear(out, sd['linear.weight'], sd['linear.bias'])                           # :360
    return torch.tanh(out) if output_mask else torch.sigmoid(out)                             # :361-364


def spectrogram_patcher_forward(sd, x, c, ns=0.01):
    """Spectrogram_Patcher.forward: model/model.py:525-549 (emb1 conditions both dense blocks, emb2 the GRU input and the
    appended channels; every layer is per-frame, so no padding mode is involved)."""
    e1, e2 = sd['emb1.weight'][c], sd['emb2.weight'][c]
    out = frame_linear(x, sd['input_layer.weight'], sd['input_layer.bias'])                    # :533-534
    out = _dec_dense_block(out, sd, ['dense1', 'dense2'], e1, ns)                             # :536
    out = _dec_dense_block(out, sd, ['dense3', 'dense4'], e1, ns)                             # :537
    rnn = bi_gru(out + e2.unsqueeze(2), sd)                                                   # :538-541
    out = torch.cat([out, rnn, e2.unsqueeze(2).expand(-1, -1, out.shape[2])], dim=1)          # :542-543
    out = F.leaky_relu(frame_linear(out, sd['dense5.weight'], sd['dense5.bias']), ns)         # :544-545
    return torch.sigmoid(frame_linear(out, sd['linear.weight'], sd['linear.bias']))           # :546-548


def enhanced_generator_forward(sd, x, c, ns=0.01, seg_len=128):
    """Enhanced_Generator.forward: model/model.py:499-502 - Encoder(enc_mode='continues') -> Decoder."""
    enc_sd = {k[len('Encoder.'):]: v for k, v in sd.items() if k.startswith('Encoder.')}
    dec_sd = {k[len('Decoder.'):]: v for k, v in sd.items() if k.startswith('Decoder.')}
    act, _, _ = encoder_forward(enc_sd, x, None, ns, seg_len, 'continues')
    return decoder_forward(dec_sd, act, c, ns, seg_len)


def combine_generator(x_dec, act, c, gen_sd, g_mode, shift, ns=0.01, seg_len=128):
    """The patcher combine rules shared by Trainer.test_step (trainer.py:206-213) and Trainer.gen_step (:274-281)."""
    if g_mode == 'naive':                                     # :206-207 / :274-275
        return x_dec + decoder_forward(gen_sd, act, c, ns, seg_len)
    if g_mode == 'targeted':                                  # :208-209 / :276-277
        return x_dec + decoder_forward(gen_sd, act, c - shift, ns, seg_len)
    if g_mode == 'targeted_residual':                         # :210-211 / :278-279
        return x_dec + x_dec * decoder_forward(gen_sd, act, c - shift, ns, seg_len, output_mask=True)
    if g_mode == 'enhanced':                                  # :212-213 / :280-281
        return x_dec + enhanced_generator_forward(gen_sd, x_dec, c - shift, ns, seg_len)
    if g_mode == 'spectrogram':
        return x_dec + spectrogram_patcher_forward(gen_sd, x_dec, c - shift, ns)
    raise NotImplementedError(g_mode)


def test_step(enc_sd, dec_sd, x, c, uniform, ns=0.01, seg_len=128, enc_mode='one_hot',
              gen_sd=None, g_mode='targeted', shift=100):
    """Trainer.test_step: trainer.py:194-221 (enc_only when gen_sd is None)."""
    act, logits, ids = encoder_forward(enc_sd, x, uniform, ns, seg_len, enc_mode)
    x_dec = decoder_forward(dec_sd, act, c, ns, seg_len)
    if gen_sd is not None:
        x_dec = combine_generator(x_dec, act, c, gen_sd, g_mode, shift, ns, seg_len)
    return x_dec, act, logits, ids


def gen_step(dec_sd, gen_sd, enc_act, c, g_mode='targeted', shift=100, ns=0.01, seg_len=128):
    """Trainer.gen_step: trainer.py:272-284 - x_dec = Decoder(enc, c) combined with the Generator's output."""
    x_dec = decoder_forward(dec_sd, enc_act, c, ns, seg_len)
    return combine_generator(x_dec, enc_act, c, gen_sd, g_mode, shift, ns, seg_len)


# ----------------------------------------------------------------------------
# pretrain_AE step: trainer.py:321-332, utils.py:50-55
# ----------------------------------------------------------------------------
def dropout_mask_shapes(B, T, c_h2=512):
    """Shapes of the six Dropout inputs of Encoder.forward (drop1..drop6), model/model.py:447-453."""
    t = [T, (T + 1) // 2, ((T + 1) // 2 + 1) // 2, (((T + 1) // 2 + 1) // 2 + 1) // 2]
    return [(B, c_h2, t[0]), (B, c_h2, t[1]), (B, c_h2, t[2]), (B, c_h2, t[3]), (B, c_h2, t[3]), (B, c_h2, t[3])]


def ae_loss_and_grads(enc_sd, dec_sd, x, c, uniform, keep_masks=None, dp=0.5, ns=0.01, seg_len=128,
                      enc_mode='one_hot'):
    """encode_step -> decode_step -> L1 loss -> backward (trainer.py:325-329) by autograd over this
    restatement.  Returns (loss, enc_grads, dec_grads, x_dec, unit_ids); grads are dicts keyed like the
    state_dicts (every parameter of both networks receives a gradient)."""
    enc_p = {k: v.detach().clone().requires_grad_(True) for k, v in enc_sd.items()}
    dec_p = {k: v.detach().clone().requires_grad_(True) for k, v in dec_sd.items()}
    act, _, ids = encoder_forward(enc_p, x, uniform, ns, seg_len, enc_mode, keep_masks=keep_masks, dp=dp)
    x_dec = decoder_forward(dec_p, act, c, ns, seg_len)
    loss = torch.mean(torch.abs(x_dec - x))                   # trainer.py:327
    loss.backward()                                           # :329
    g_enc = {k: v.grad if v.grad is not None else torch.zeros_like(v) for k, v in enc_p.items()}
    g_dec = {k: v.grad if v.grad is not None else torch.zeros_like(v) for k, v in dec_p.items()}
    return loss.detach(), g_enc, g_dec, x_dec.detach(), ids


def clip_grad_norm(grads, max_norm):
    """nn.utils.clip_grad_norm_ over ONE network (utils.py:53-55): returns (total_norm, scaled grads)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, {k: g * coef for k, g in grads.items()}


def adam_step(params, grads, state, lr=1e-4, betas=(0.5, 0.9), eps=1e-8):
    """torch.optim.Adam (trainer.py:64-66: lr, betas=(0.5, 0.9), no weight decay, no amsgrad), one step in place
    on `params` / `state` ({'step': int, 'm': {...}, 'v': {...}})."""
    state['step'] = state.get('step', 0) + 1
    t = state['step']
    m, v = state.setdefault('m', {}), state.setdefault('v', {})
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k]
        m[k] = b1 * m.get(k, torch.zeros_like(p)) + (1 - b1) * g
        v[k] = b2 * v.get(k, torch.zeros_like(p)) + (1 - b2) * g * g
        denom = (v[k].sqrt() / math.sqrt(1 - b2 ** t)) + eps
        p.sub_((lr / (1 - b1 ** t)) * m[k] / denom)
    return params


def pretrain_ae_step(enc_sd, dec_sd, opt_state, x, c, uniform, keep_masks=None, dp=0.5, ns=0.01, seg_len=128,
                     enc_mode='one_hot', lr=1e-4, max_grad_norm=5.0):
    """One iteration of Trainer.train(mode='pretrain_AE'): trainer.py:321-332.  Updates enc_sd / dec_sd in place
    (one Adam over both networks' parameters, per-network clipping) and returns (loss, norm_enc, norm_dec)."""
    loss, g_enc, g_dec, _, _ = ae_loss_and_grads(enc_sd, dec_sd, x, c, uniform, keep_masks, dp, ns, seg_len, enc_mode)
    n_enc, g_enc = clip_grad_norm(g_enc, max_grad_norm)       # utils.py:53-55, one norm per network
    n_dec, g_dec = clip_grad_norm(g_dec, max_grad_norm)
    params = {('e', k): v for k, v in enc_sd.items()}
    params.update({('d', k): v for k, v in dec_sd.items()})
    grads = {('e', k): v for k, v in g_enc.items()}
    grads.update({('d', k): v for k, v in g_dec.items()})
    adam_step(params, grads, opt_state, lr=lr)
    return loss, n_enc, n_dec


# ----------------------------------------------------------------------------
# Driver glue: convert.py:36, 120-221
# ----------------------------------------------------------------------------
MIN_LEN = 9  # convert.py:36


def segment_plan(n_frames, seg_len):
    """Chunking rule of convert()/encode(): convert.py:139-165, 189-213.

    Returns (padded_len, [(start, stop), ...], keep_units) where the model is
    called once per (start, stop) slice of the (zero-padded to MIN_LEN) utterance
    and `keep_units` is the number of unit frames kept from a padded utterance
    (None = all)."""
    padded = max(n_frames, MIN_LEN)                           # :140-143
    keep = MIN_LEN // 8 if n_frames < MIN_LEN else None       # :147-148
    if padded <= seg_len:                                     # :145
        return padded, [(0, padded)], keep
    plan = []
    for idx in range(0, padded, seg_len):                     # :153
        if idx + 2 * seg_len > padded:                        # :154-155 tail drops the last frame
            start, stop = idx, padded - 1
        else:
            start, stop = idx, idx + seg_len
        if stop - start >= seg_len:                           # :159
            plan.append((start, stop))
        elif idx == 0:
            raise RuntimeError('Please check if input is too short!')
    return padded, plan, None


def format_encodings(encodings):
    """write_encodings(): convert.py:120-126 - one line per unit frame, ints, space separated."""
    lines = []
    for enc in encodings:
        lines.append(' '.join(str(int(e)) for e in enc))
    return '\n'.join(lines) + ('\n' if lines else '')
