"""Helpers shared by the -m gpu tests: drive the C-ABI building blocks from torch tensors."""
import ctypes as C

import torch
import torch.nn.functional as F

import zs_b200  # noqa: F401
from zs_b200 import _lib


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up(v, m):
    return (v + m - 1) // m * m


def pack_weight(W, operand_dtype, ps=False):
    """(C_out, C_in, k) fp32 -> [m_rows][k][c_in_pad] operand dtype (+ optional pixel-shuffle row permutation)."""
    C_out, C_in, k = W.shape
    m_rows, c_pad = round_up(C_out, 128), round_up(C_in, 64)
    Wp = torch.zeros(m_rows, k, c_pad, dtype=operand_dtype, device=W.device)
    rows = torch.arange(C_out, device=W.device)
    if ps:
        c, r = rows // 2, rows % 2
        rows = (c // 64) * 128 + r * 64 + c % 64
    Wp[rows, :, :C_in] = W.permute(0, 2, 1).to(operand_dtype)
    return Wp.contiguous(), m_rows, c_pad


def pack_act(x, halo, operand='fp16', lrelu=False, ns=0.01):
    """(B, C, T) fp32 cuda -> channels-last operand buffer with reflected halo, via zs_pack_nct."""
    B, Cc, T = x.shape
    rows, pitch = round_up(T + 2 * halo, 2), round_up(Cc, 8)
    dt = torch.float16 if operand == 'fp16' else torch.bfloat16
    buf = torch.full((B, rows, pitch), float('nan'), dtype=dt, device=x.device)
    _lib.check(_lib.lib().zs_pack_nct(ptr(x), B, Cc, T, ptr(buf), rows, pitch, halo, 0, int(lrelu), ns,
                                      _lib.OPERANDS[operand], 0, stream()))
    return buf, rows, pitch


def conv_cl(x, W, bias, *, stride=1, lrelu=False, inorm=False, ns=0.01, operand='fp16', nb_hint=0, act=0):
    """Run zs_conv1d_cl on an fp32 (B, C, T) input and return the fp32 (B, C_out, T_out) output."""
    B, C_in, T = x.shape
    C_out, _, k = W.shape
    dt = torch.float16 if operand == 'fp16' else torch.bfloat16
    halo = k // 2
    buf, rows, pitch = pack_act(x, halo, operand)
    Wp, m_rows, c_pad = pack_weight(W, dt)
    T_out = (T + stride - 1) // stride
    bias_p = torch.zeros(m_rows, dtype=torch.float32, device=x.device)
    bias_p[:C_out] = bias
    out = torch.full((B, C_out, T_out), float('nan'), dtype=torch.float32, device=x.device)
    d = _lib.ConvDesc()
    d.w, d.m_rows, d.m_valid, d.taps, d.c_in_pad, d.w_taps, d.bank = Wp.data_ptr(), m_rows, C_out, k, c_pad, k, 0
    d.in_, d.in_rows, d.in_pitch, d.in_row0, d.c_in_valid = buf.data_ptr(), rows, pitch, 0, C_in
    d.stride, d.B, d.T_out = stride, B, T_out
    d.bias, d.spk, d.n_spk, d.lrelu, d.ns, d.inorm = bias_p.data_ptr(), None, 1, int(lrelu), ns, int(inorm)
    d.res_mode, d.res = 0, None
    d.act, d.out_mode = act, 2
    d.out, d.out_rows, d.out_pitch, d.out_halo, d.out_choff = out.data_ptr(), 0, 0, 0, 0
    d.accumulate, d.operand, d.nb_hint = 0, _lib.OPERANDS[operand], nb_hint
    _lib.check(_lib.lib().zs_conv1d_cl(C.byref(d), stream()))
    torch.cuda.synchronize()
    return out


def conv_ref(x, W, bias, *, stride=1, lrelu=False, inorm=False, ns=0.01, operand='fp16'):
    """Same layer in fp64 on the operand-rounded inputs (what the tensor cores multiply)."""
    dt = torch.float16 if operand == 'fp16' else torch.bfloat16
    xr, Wr = x.to(dt).double(), W.to(dt).double()
    k = W.shape[2]
    pad = (k // 2, k // 2 - 1) if k % 2 == 0 else (k // 2, k // 2)
    y = F.conv1d(F.pad(xr, pad, mode='reflect') if k > 1 else xr, Wr, bias.double(), stride=stride)
    if lrelu:
        y = F.leaky_relu(y, ns)
    if inorm:
        mu = y.mean(2, keepdim=True)
        var = ((y - mu) ** 2).mean(2, keepdim=True)
        y = (y - mu) / torch.sqrt(var + 1e-5)
    return y.float()


def conv_cl_to_cl(x, W, bias, *, lrelu=False, inorm=False, ns=0.01, operand='fp16', nb_hint=0, halo_out=1):
    """zs_conv1d_cl with the channels-last operand output (the staged / TMA-store epilogue); returns fp32 (B, C_out, T)."""
    B, C_in, T = x.shape
    C_out, _, k = W.shape
    dt = torch.float16 if operand == 'fp16' else torch.bfloat16
    buf, rows, pitch = pack_act(x, k // 2, operand)
    Wp, m_rows, c_pad = pack_weight(W, dt)
    bias_p = torch.zeros(m_rows, dtype=torch.float32, device=x.device)
    bias_p[:C_out] = bias
    out_rows, out_pitch = round_up(T + 2 * halo_out, 2), round_up(C_out, 8)
    out = torch.full((B, out_rows, out_pitch), float('nan'), dtype=dt, device=x.device)
    d = _lib.ConvDesc()
    d.w, d.m_rows, d.m_valid, d.taps, d.c_in_pad, d.w_taps, d.bank = Wp.data_ptr(), m_rows, C_out, k, c_pad, k, 0
    d.in_, d.in_rows, d.in_pitch, d.in_row0, d.c_in_valid = buf.data_ptr(), rows, pitch, 0, C_in
    d.stride, d.B, d.T_out = 1, B, T
    d.bias, d.spk, d.n_spk, d.lrelu, d.ns, d.inorm = bias_p.data_ptr(), None, 1, int(lrelu), ns, int(inorm)
    d.res_mode, d.res = 0, None
    d.act, d.out_mode = 0, 0
    d.out, d.out_rows, d.out_pitch, d.out_halo, d.out_choff = out.data_ptr(), out_rows, out_pitch, halo_out, 0
    d.accumulate, d.operand, d.nb_hint = 0, _lib.OPERANDS[operand], nb_hint
    _lib.check(_lib.lib().zs_conv1d_cl(C.byref(d), stream()))
    torch.cuda.synchronize()
    body = out[:, halo_out:halo_out + T, :C_out].float().permute(0, 2, 1)
    # reflected halo rows
    for h in range(1, halo_out + 1):
        assert torch.equal(out[:, halo_out - h, :C_out], out[:, halo_out + h, :C_out])
        assert torch.equal(out[:, halo_out + T - 1 + h, :C_out], out[:, halo_out + T - 1 - h, :C_out])
    return body
