"""The pretrain_AE iteration of the reference (`Trainer.train(mode='pretrain_AE')`, trainer.py:321-332) on the
B200 path:

    enc_act, enc = Encoder(x) ; x_dec = Decoder(enc_act, c) ; loss = mean|x_dec - x|
    zero_grad ; backward ; clip_grad_norm_(5) per network ; Adam(lr 1e-4, betas (0.5, 0.9)).step()

`PretrainAE.step` runs all of it through libzsae.so (training forward, fused L1 loss + backward, fused clip+Adam,
in-place re-pack of the tensor-core operands); PyTorch only owns the memory, the streams and - with more than one
rank - the gradient all-reduce (`torch.distributed`, NCCL over NVLink; gloo in the CPU tests).  The decoder's
gradients are reduced on a side stream while the encoder's backward still runs.

`encode_step` / `decode_step` wrap the same kernels as `torch.autograd.Function`s for callers that build the loss
themselves and call `loss.backward()` (the reference's own loop shape, trainer.py:325-329).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .model import Decoder, Encoder, _ptr, _stream, gumbel_from_uniform


def flatten_parameters(module):
    """Moves every parameter of `module` into ONE contiguous fp32 buffer (parameters become views of it, names and
    shapes unchanged; every tensor starts on a 256-byte boundary, the zero gaps are inert for the norm and Adam) and
    returns (flat_params, {name: (offset, numel)})."""
    params = list(module.named_parameters())
    dev = params[0][1].device
    align = 64
    total = sum((p.numel() + align - 1) // align * align for _, p in params)
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    layout, off = {}, 0
    for name, p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view_as(p)
        layout[name] = (off, n)
        off += (n + align - 1) // align * align
    return flat, layout


def views_like(flat, layout, module):
    shapes = {n: p.shape for n, p in module.named_parameters()}
    return {n: flat[o:o + k].view(shapes[n]) for n, (o, k) in layout.items()}


def reduce_gradients(flat_grad, group=None):
    """The one collective of the training path: summing all-reduce of a network's flat fp32 gradient
    over the data-parallel ranks (NCCL over NVLink on the GPUs, gloo in the CPU tests).  The sum is divided by the
    world size where it is consumed (zs_adam_step's `grad_scale`), which makes the update that of the reference
    step on the concatenated batch: mean-L1 over equal per-rank batches, no cross-sample statistics anywhere."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


class _Net:
    """Flat parameter / gradient / Adam-moment storage of one network."""

    def __init__(self, module):
        self.module = module
        self.flat, self.layout = flatten_parameters(module)
        self.grad = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.grad_views = views_like(self.grad, self.layout, module)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=self.flat.device)


class PretrainAE:
    """One object = the (Encoder, Decoder, ae_opt) triple of trainer.py:58-66 with `step()` = trainer.py:321-332."""

    def __init__(self, encoder: Encoder, decoder: Decoder, lr=1e-4, betas=(0.5, 0.9), eps=1e-8, max_grad_norm=5.0,
                 loss_scale=None, process_group=None, cpu_noise=False, use_graph=True, async_wgrad=True):
        if encoder.enc_mode != 'one_hot':
            raise RuntimeError("PretrainAE: the training path implements enc_mode 'one_hot'")
        self.enc, self.dec = _Net(encoder), _Net(decoder)
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.loss_scale = loss_scale            # None: 2^15 * B at the first step
        self.step_count = 0
        self.good_steps = 0
        self.cpu_noise = cpu_noise
        # weight-gradient GEMMs on the library's side streams (include/zs_ae.h zs_wgrad_async): single-rank steps only.
        # A 2-GPU run with it enabled next to the captured NCCL all-reduces did not finish within its time limit
        # (gpurun_out/r4d, cause not isolated), so data-parallel steps keep the in-order launches measured in round 2.
        self.async_wgrad = bool(async_wgrad)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.async_wgrad = async_wgrad == 'force' or (self.async_wgrad and self.world == 1)     # 'force': tools/wgrad_async_dp_probe.py
        dev = self.enc.flat.device
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=dev)
        self._skipped_host = torch.zeros(1, dtype=torch.int32).pin_memory() if dev.type == 'cuda' else None
        self._skip_event = None
        self._skip_ring = []
        self._skip_hosts = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(4)] if dev.type == 'cuda' else []
        self.side = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.n_skipped = 0
        # CUDA-graph replay of the whole iteration (single rank; ~190 launches otherwise cost more host time than the
        # GPU needs to run them).  What changes between replays (dropout seed, Adam's bias corrections) lives in
        # `_meta` and is advanced by kernels inside the graph.
        # With more than one rank the two NCCL all-reduces are captured too (NCCL collectives are capturable; the side
        # stream forks from and joins the capturing stream), so data-parallel steps replay as one graph as well.
        self.use_graph = bool(use_graph) and not cpu_noise and dev.type == 'cuda'
        self._graph = None
        self._graph_key = None
        self._warm = 0
        # step state in DEVICE memory (include/zs_ae.h zs_train_meta_*): dropout seed, Adam bias corrections, the
        # all-networks-or-none apply flag, the count of applied steps.  It advances on the stream - a graph replay
        # needs no host write that an earlier, still running replay could observe too late.
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.seed_salt = (0x1234567 + 0xD1B54A32D192ED03 * self.rank) & (2 ** 64 - 1)     # ranks draw different masks
        self._meta = torch.zeros(8, dtype=torch.int32, device=dev)

    # ---- pieces -----------------------------------------------------------------------------------------
    def _noise(self, B, T8, dev):
        shape = (B, T8, self.enc.module.enc_size)
        if self.cpu_noise:      # the reference's RNG contract (model/model.py:96: torch.rand on the CPU generator)
            return gumbel_from_uniform(torch.rand(shape)).to(dev, non_blocking=True)
        return gumbel_from_uniform(torch.rand(shape, device=dev))

    def _poll_skip(self):
        """Dynamic loss scaling without a host sync on the step just issued: halve after an overflow, double after 1000
        clean steps.  One rank: the flag of the last finished step is read whenever it has arrived.  Several ranks: every
        rank must change the scale (and re-capture its graph) at the SAME iteration, so the flag of iteration k-2 is
        awaited at iteration k - deterministic, and two iterations stay in flight."""
        if self.world > 1:
            if len(self._skip_ring) < 2:
                return
            ev, host = self._skip_ring.pop(0)
            ev.synchronize()
            flag = int(host[0])
        else:
            if self._skip_event is None or not self._skip_event.query():
                return
            flag = int(self._skipped_host[0])
            self._skip_event = None
        if flag:
            self.loss_scale = max(self.loss_scale * 0.5, 1.0)
            self.n_skipped += 1
            self.good_steps = 0
            self.skipped.zero_()
        else:
            self.good_steps += 1
            if self.good_steps >= 1000 and self.loss_scale < 2.0 ** 24:
                self.loss_scale *= 2.0
                self.good_steps = 0

    def _allreduce(self, net, stream):
        """Summing all-reduce of one network's flat gradient on `stream` (zs_adam_step divides by the world size)."""
        with torch.cuda.stream(stream):
            reduce_gradients(net.grad, self.pg)

    def _meta_ptr(self, word):
        return C.c_void_p(self._meta.data_ptr() + 4 * word)

    def _optim(self):
        """utils.py:53-55 per-network clip + trainer.py:332 ae_opt.step(): ONE Adam over both networks, so an fp16
        overflow in either skips both (and does not count as a step)."""
        lib = _lib.lib()
        for net in (self.enc, self.dec):
            net.sqnorm.zero_()
            _lib.check(lib.zs_grad_sqnorm(_ptr(net.grad), net.flat.numel(), _ptr(net.sqnorm), _stream()))
        _lib.check(lib.zs_train_meta_commit(self._meta_ptr(0), _ptr(self.enc.sqnorm), _ptr(self.dec.sqnorm), self.betas[0],
                                            self.betas[1], _ptr(self.skipped), _stream()))
        for net in (self.enc, self.dec):
            _lib.check(lib.zs_adam_step(_ptr(net.flat), _ptr(net.grad), _ptr(net.m), _ptr(net.v), net.flat.numel(),
                                        _ptr(net.sqnorm), 1.0 / self.world, self.max_grad_norm, self.lr, self.betas[0],
                                        self.betas[1], self.eps, 1, self._meta_ptr(2), _ptr(self.skipped), _stream()))

    # ---- the iteration ----------------------------------------------------------------------------------
    def forward_backward(self, x, c, noise=None, dropout_seed=None, keep_masks=None):
        """encode_step -> decode_step -> L1 -> backward.  Leaves the (local) gradients in the flat buffers and
        returns (loss (device scalar), unit ids).  The dropout masks come from the device-resident seed (a new one
        every iteration, different on every rank) unless `dropout_seed` / `keep_masks` pin them."""
        enc, dec = self.enc.module, self.dec.module
        B, _, T = x.shape
        dev = x.device
        if self.loss_scale is None:
            self.loss_scale = float(2 ** 15 * B)
        if noise is None:
            noise = self._noise(B, enc.t8(T), dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().zs_train_meta_begin(self._meta_ptr(0), self.seed_salt, _stream()))
        seed_dev = None
        if dropout_seed is None:
            dropout_seed, seed_dev = 0, self._meta[0:2].view(torch.int64)
        self.enc.grad.zero_()
        self.dec.grad.zero_()
        self.loss.zero_()
        act, _, ids = enc.forward_train(x, noise, dropout_seed, keep_masks, seed_dev)   # trainer.py:325
        dec.forward_train(act, c)                                                        # :326
        lib = _lib.lib()
        with torch.cuda.device(dev):
            if self.async_wgrad:            # weight-gradient GEMMs leave the data-gradient chain (include/zs_ae.h)
                _lib.check(lib.zs_wgrad_async(1))
            try:
                d_act = dec.backward(self.dec.grad_views, self.loss_scale, target=x, loss_out=self.loss)   # :327-329
                if self.world > 1:          # decoder gradients travel while the encoder backward runs
                    _lib.check(lib.zs_wgrad_join(_stream()))          # (no-op unless async_wgrad was forced on)
                    self.side.wait_stream(torch.cuda.current_stream())
                    self._allreduce(self.dec, self.side)
                enc.backward(d_act, self.enc.grad_views, self.loss_scale, d_act_scale=self.loss_scale)
            finally:
                _lib.check(lib.zs_wgrad_join(_stream()))
                _lib.check(lib.zs_wgrad_async(0))
        return self.loss, ids

    def _finish(self, dev):
        """Encoder gradient reduce (the decoder's is already in flight on the side stream), then trainer.py:330-332:
        per-network clip, ONE Adam over both networks, and the in-place refresh of the tensor-core operands."""
        if self.world > 1:
            self.side.wait_stream(torch.cuda.current_stream())
            self._allreduce(self.enc, self.side)
            torch.cuda.current_stream().wait_stream(self.side)
        with torch.cuda.device(dev):
            self._optim()
            self.enc.module._repack(dev)
            self.dec.module._repack(dev)

    def _graph_body(self):
        """The iteration on static buffers; everything here is captured into the CUDA graph."""
        self.forward_backward(self._x_static, self._c_static)
        self._finish(self._x_static.device)

    def _step_graph(self, x, c):
        key = (tuple(x.shape), float(self.loss_scale) if self.loss_scale else None)
        if self._graph is None or key != self._graph_key:
            self._x_static = torch.empty_like(x)
            self._c_static = torch.empty_like(c)
            self._graph = torch.cuda.CUDAGraph()
            self._x_static.copy_(x)
            self._c_static.copy_(c)
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.pg)         # every rank captures the same collectives at the same step
            with torch.cuda.graph(self._graph):     # (a capture pass does not execute: `_meta` is untouched)
                self._graph_body()
            self._graph_key = key
        self._x_static.copy_(x, non_blocking=True)
        self._c_static.copy_(c, non_blocking=True)
        self._graph.replay()
        self.enc.module._packed_key = self.dec.module._packed_key = None
        return self.loss

    def step(self, x, c, noise=None, dropout_seed=None, keep_masks=None):
        """One pretrain_AE iteration (trainer.py:321-332).  Returns the loss as a device tensor (no host sync)."""
        self._poll_skip()
        self.step_count += 1
        plain = noise is not None or dropout_seed is not None or keep_masks is not None
        if self.use_graph and not plain and self._warm >= 2:     # two eager steps first: allocations, attributes, loss scale
            loss = self._step_graph(x, c)
            self._after_step()
            return loss
        self._warm += 1
        loss, _ = self.forward_backward(x, c, noise, dropout_seed, keep_masks)
        self._finish(x.device)
        self.enc.module._packed_key = self.dec.module._packed_key = None     # the eval handles are stale now
        self._after_step()
        return loss

    def _after_step(self):
        if self._skipped_host is None:
            return
        if self.world > 1:
            host = self._skip_hosts[self.step_count % len(self._skip_hosts)]
            host.copy_(self.skipped, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._skip_ring.append((ev, host))
        elif self._skip_event is None:
            self._skipped_host.copy_(self.skipped, non_blocking=True)
            self._skip_event = torch.cuda.Event()
            self._skip_event.record()

    def applied_steps(self):
        """Optimiser steps actually applied (overflow-skipped iterations do not count); synchronises."""
        return int(self._meta[5].item())

    def grad_norms(self):
        """(encoder, decoder) gradient L2 norms of the last step, before clipping (host floats; synchronises)."""
        return (float(self.enc.sqnorm.sqrt()) / self.world, float(self.dec.sqnorm.sqrt()) / self.world)


# ---------------------------------------------------------------------------------------------------------
# autograd wrappers: the reference's loop shape (trainer.py:246-254, 325-329) with loss.backward()
# ---------------------------------------------------------------------------------------------------------
AUTOGRAD_LOSS_SCALE = 2.0 ** 20     # fallback when an encoder backward runs without a decoder backward before it
_last_scale = [None]                 # loss scale the decoder backward of the running loss.backward() chose


def _pick_scale(d):
    """Power-of-two loss scale that puts max|d| near 0.5 in the fp16 gradient activations (a mean-L1 loss over
    B x 513 x 128 elements gives 2^20 at B = 32).  Raises on non-finite incoming gradients."""
    amax = float(d.abs().max())
    if not (amax == amax) or amax == float('inf'):
        raise RuntimeError('autograd wrapper: non-finite incoming gradient')
    if amax == 0.0:
        return AUTOGRAD_LOSS_SCALE
    import math
    return float(2.0 ** max(0, min(40, math.floor(math.log2(0.5 / amax)))))


def _flat_grads(m, dev):
    names = [n for n, _ in m.named_parameters()]
    params = [p for _, p in m.named_parameters()]
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    grads, off = {}, 0
    for n, p in zip(names, params):
        grads[n] = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return names, flat, grads


def _check_ctx(ctx, what):
    if ctx.module._train_ctx_version != ctx.version:
        raise RuntimeError(f'{what}: the module ran another training forward before this backward - it keeps ONE set of '
                           'saved activations (call backward before the next forward of the same module)')


class _EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, noise, dropout_seed, seed_dev, *params):
        act, logits, _ = module.forward_train(x, noise, dropout_seed, seed_dev=seed_dev)
        ctx.module, ctx.version = module, module._train_ctx_version
        ctx.mark_non_differentiable(logits)
        return act, logits

    @staticmethod
    def backward(ctx, d_act, _d_logits):
        m = ctx.module
        _check_ctx(ctx, 'encode_step')
        names, flat, grads = _flat_grads(m, d_act.device)
        scale = _last_scale[0] or _pick_scale(d_act) / 64.0
        m.backward(d_act, grads, scale, d_act_scale=1.0)
        if not bool(torch.isfinite(flat).all()):
            raise RuntimeError(f'encode_step backward: fp16 gradient overflow at loss scale {scale:g}')
        return (None, None, None, None, None) + tuple(grads[n] for n in names)


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, c, *params):
        out = module.forward_train(x, c)
        ctx.module, ctx.version = module, module._train_ctx_version
        return out

    @staticmethod
    def backward(ctx, d_spec):
        m = ctx.module
        _check_ctx(ctx, 'decode_step')
        names, flat, grads = _flat_grads(m, d_spec.device)
        scale = _pick_scale(d_spec)
        _last_scale[0] = scale
        d_act = m.backward(grads, scale, d_spec=d_spec)
        if not bool(torch.isfinite(flat).all()) or not bool(torch.isfinite(d_act).all()):
            raise RuntimeError(f'decode_step backward: fp16 gradient overflow at loss scale {scale:g}')
        return (None, d_act / scale, None) + tuple(grads[n] for n in names)


def encode_step(encoder, x, noise=None, dropout_seed=None):
    """`Trainer.encode_step` (trainer.py:246-249) with an autograd graph: returns (enc_act, enc).

    Every call draws NEW dropout masks (the reference's nn.Dropout does): the seed of the counter-based masks comes from
    torch's CUDA generator - the generator the reference's dropout consumes with the model on a GPU - so
    `torch.manual_seed` pins it and the CPU generator is consumed by the Gumbel draw alone, as in the reference.
    Only `enc_act` carries a gradient: `enc` (the logits) is returned for inspection, a loss term on it contributes
    nothing (the reference's pretrain_AE loss, trainer.py:327, uses x_dec only)."""
    if noise is None:
        B, _, T = x.shape
        noise = gumbel_from_uniform(torch.rand(B, encoder.t8(T), encoder.enc_size)).to(x.device)
    seed_dev = None
    if dropout_seed is None:
        seed_dev = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=x.device)
        dropout_seed = 0
    _last_scale[0] = None
    return _EncodeFn.apply(encoder, x, noise, dropout_seed, seed_dev, *encoder.parameters())


def decode_step(decoder, enc_act, c):
    """`Trainer.decode_step` (trainer.py:251-254) with an autograd graph."""
    return _DecodeFn.apply(decoder, enc_act, c, *decoder.parameters())


class TrainerSteps:
    """The Trainer's train-mode step wrappers for this path (trainer.py:238-254, 272-284) on the B200 modules, with autograd
    graphs, so the reference's training loops can call them unchanged:

        C, X = steps.permute_data(next(loader))      # trainer.py:238-244
        enc_act, enc = steps.encode_step(X)          # :246-249
        x_dec = steps.decode_step(enc_act, C)        # :251-254
        x_gen = steps.gen_step(enc_act, C)           # :272-284  (Decoder output combined with the Generator's)

    `gen_step` differentiates through a `Decoder`-type Generator (g_mode naive / targeted / targeted_residual); the
    alternate patchers (enhanced / spectrogram) are inference-only here - their training is the stage-2 GAN loop."""

    def __init__(self, encoder, decoder, generator=None, g_mode='targeted', n_speakers=102, n_target_speakers=2):
        self.Encoder, self.Decoder, self.Generator, self.g_mode = encoder, decoder, generator, g_mode
        self.shift_c = n_speakers - n_target_speakers          # trainer.py:44-45
        self.device = next(encoder.parameters()).device

    def permute_data(self, data, load_mel=False):
        """(speaker ids (B,), lin (B, T, 513)[, mel (B, T, 80)]) from the loader -> device tensors, spectrograms permuted
        to (B, C, T) and - like utils.to_var (utils.py:43-45) - requiring grad."""
        c = data[0].to(self.device)
        x = data[1].to(self.device).float().requires_grad_(True).permute(0, 2, 1)
        if load_mel:
            return c, x, data[2].to(self.device).float().requires_grad_(True).permute(0, 2, 1)
        return c, x

    def encode_step(self, x, noise=None):
        return encode_step(self.Encoder, x, noise)

    def decode_step(self, enc, c):
        return decode_step(self.Decoder, enc, c)

    def gen_step(self, enc, c):
        x_dec = decode_step(self.Decoder, enc, c)
        if self.Generator is None:
            raise RuntimeError('gen_step needs a Generator')
        if self.g_mode == 'naive':
            return x_dec + decode_step(self.Generator, enc, c)
        if self.g_mode == 'targeted':
            return x_dec + decode_step(self.Generator, enc, c - self.shift_c)
        if self.g_mode == 'targeted_residual':
            return x_dec + x_dec * decode_step(self.Generator, enc, c - self.shift_c)
        if self.g_mode in ('enhanced', 'spectrogram'):
            if torch.is_grad_enabled() and self.Generator.training:
                raise NotImplementedError(f"gen_step: training a g_mode {self.g_mode!r} patcher is outside the autoencoder hot path")
            return x_dec + self.Generator(x_dec.detach(), c - self.shift_c)
        raise NotImplementedError('Invalid generator mode to call gen_step()!')
