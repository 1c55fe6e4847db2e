"""ORACLE - test infrastructure, NOT product code.

CPU fp32 restatement of the reference's 2-D critic / classifier forward passes (forward part):
`PatchDiscriminator.forward` (model/model.py:153-166) and `TargetClassifier.forward` (model/model.py:212-223), eval mode
(Dropout2d is the identity), as plain functions over a reference-layout ``state_dict``.

Parity pin: the live reference modules, imported in the build container by ``tests/golden/make_golden_critic.py``
(fixtures ``tests/golden/critic_*.npz``); ``tests/test_oracle_golden.py`` replays them against this file (<= 2e-5).
Nothing under ``zerospeech-tts-without-t_b200/`` may import it.
"""
import torch
import torch.nn.functional as F

IN_EPS = 1e-5     # nn.InstanceNorm2d default (model/model.py:139-144)


def conv2d_same(x, w, b, seg_len, stride):
    """pad_layer(is_2d=True) + nn.Conv2d: model/model.py:29-40 (odd k pads k//2 on all four sides)."""
    k = w.shape[2]
    pad = (k // 2, k // 2 - 1) * 2 if k % 2 == 0 else (k // 2,) * 4
    x = F.pad(x, pad, mode='constant' if seg_len < 64 else 'reflect') if k > 1 else x
    return F.conv2d(x, w, b, stride=stride)


def instance_norm2d(x):
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + IN_EPS)


def critic_trunk(sd, x, ns=0.2, seg_len=128):
    """conv_block x 6 (model/model.py:146-151, 155-160): pad -> conv -> leaky_relu -> InstanceNorm2d (-> Dropout2d, eval)."""
    out = x.unsqueeze(1)                                        # :154
    for i in range(1, 7):
        out = conv2d_same(out, sd[f'conv{i}.weight'], sd[f'conv{i}.bias'], seg_len, stride=2 if i < 6 else 1)
        out = instance_norm2d(F.leaky_relu(out, ns))
    return out


def patch_discriminator(sd, x, ns=0.2, seg_len=128, classify=True):
    """model/model.py:153-166: (mean_val (B,), logits (B, n_class))."""
    out = critic_trunk(sd, x, ns, seg_len)
    val = F.conv2d(out, sd['conv7.weight'], sd['conv7.bias'])
    mean_val = val.view(val.size(0), -1).mean(dim=1)
    if not classify:
        return mean_val
    logits = F.conv2d(out, sd['conv_classify.weight'], sd['conv_classify.bias'])
    return mean_val, logits.view(logits.size(0), -1)


def target_classifier(sd, x, ns=0.2, seg_len=128):
    """model/model.py:212-223."""
    out = critic_trunk(sd, x, ns, seg_len)
    logits = F.conv2d(out, sd['conv_classify.weight'], sd['conv_classify.bias'])
    return logits.view(logits.size(0), -1)
