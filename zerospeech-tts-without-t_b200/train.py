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
                 loss_scale=None, process_group=None, cpu_noise=False, use_graph=True):
        if encoder.enc_mode != 'one_hot':
            raise RuntimeError("PretrainAE: the training path implements enc_mode 'one_hot'")
        self.enc, self.dec = _Net(encoder), _Net(decoder)
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.loss_scale = loss_scale            # None: 2^15 * B at the first step
        self.step_count = 0
        self.good_steps = 0
        self.cpu_noise = cpu_noise
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = self.enc.flat.device
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=dev)
        self._skipped_host = torch.zeros(1, dtype=torch.int32).pin_memory() if dev.type == 'cuda' else None
        self._skip_event = None
        self.side = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.n_skipped = 0
        # CUDA-graph replay of the whole iteration (single rank; ~190 launches otherwise cost more host time than the
        # GPU needs to run them).  What changes between replays lives in device memory, refreshed from `_meta_host`
        # by a copy node at the head of the graph: the dropout seed and Adam's two bias corrections.
        self.use_graph = bool(use_graph) and self.world == 1 and not cpu_noise and dev.type == 'cuda'
        self._graph = None
        self._graph_key = None
        self._warm = 0
        if dev.type == 'cuda':
            self._meta_host = torch.zeros(4, dtype=torch.float32).pin_memory()      # [seed lo, seed hi (as raw bits), bc1, bc2]
            self._meta_dev = torch.zeros(4, dtype=torch.float32, device=dev)

    # ---- pieces -----------------------------------------------------------------------------------------
    def _noise(self, B, T8, dev):
        shape = (B, T8, self.enc.module.enc_size)
        if self.cpu_noise:      # the reference's RNG contract (model/model.py:96: torch.rand on the CPU generator)
            return gumbel_from_uniform(torch.rand(shape)).to(dev, non_blocking=True)
        return gumbel_from_uniform(torch.rand(shape, device=dev))

    def _poll_skip(self):
        """Dynamic loss scaling, one step late (no host sync in the step): halve after an overflow, double after
        1000 clean steps."""
        if self._skip_event is not None and self._skip_event.query():
            if int(self._skipped_host[0]):
                self.loss_scale = max(self.loss_scale * 0.5, 1.0)
                self.n_skipped += 1
                self.good_steps = 0
                self.skipped.zero_()
            else:
                self.good_steps += 1
                if self.good_steps >= 1000 and self.loss_scale < 2.0 ** 24:
                    self.loss_scale *= 2.0
                    self.good_steps = 0
            self._skip_event = None

    def _allreduce(self, net, stream):
        """Summing all-reduce of one network's flat gradient on `stream` (zs_adam_step divides by the world size)."""
        with torch.cuda.stream(stream):
            reduce_gradients(net.grad, self.pg)

    def _optim(self, net, bc_dev=None):
        lib = _lib.lib()
        n = net.flat.numel()
        net.sqnorm.zero_()
        _lib.check(lib.zs_grad_sqnorm(_ptr(net.grad), n, _ptr(net.sqnorm), _stream()))
        _lib.check(lib.zs_adam_step(_ptr(net.flat), _ptr(net.grad), _ptr(net.m), _ptr(net.v), n, _ptr(net.sqnorm),
                                    1.0 / self.world, self.max_grad_norm, self.lr, self.betas[0], self.betas[1], self.eps,
                                    max(self.step_count, 1), _ptr(bc_dev), _ptr(self.skipped), _stream()))

    # ---- the iteration ----------------------------------------------------------------------------------
    def forward_backward(self, x, c, noise=None, dropout_seed=None, keep_masks=None, seed_dev=None):
        """encode_step -> decode_step -> L1 -> backward.  Leaves the (local) gradients in the flat buffers and
        returns (loss (device scalar), unit ids)."""
        enc, dec = self.enc.module, self.dec.module
        B, _, T = x.shape
        dev = x.device
        if self.loss_scale is None:
            self.loss_scale = float(2 ** 15 * B)
        if noise is None:
            noise = self._noise(B, enc.t8(T), dev)
        if dropout_seed is None:
            dropout_seed = self.step_count * 0x9E3779B1 + 12345
        self.enc.grad.zero_()
        self.dec.grad.zero_()
        self.loss.zero_()
        act, _, ids = enc.forward_train(x, noise, dropout_seed, keep_masks, seed_dev)   # trainer.py:325
        dec.forward_train(act, c)                                                        # :326
        d_act = dec.backward(self.dec.grad_views, self.loss_scale, target=x, loss_out=self.loss)   # :327-329
        if self.world > 1:                  # decoder gradients travel while the encoder backward runs
            self.side.wait_stream(torch.cuda.current_stream())
            self._allreduce(self.dec, self.side)
        enc.backward(d_act, self.enc.grad_views, self.loss_scale, d_act_scale=self.loss_scale)
        return self.loss, ids

    def _set_meta(self):
        import struct
        seed = (self.step_count * 0x9E3779B97F4A7C15 + 0x1234567) & (2 ** 64 - 1)
        # NaN bit patterns would not survive a float round trip through python: write the raw words instead
        self._meta_host.view(torch.int32)[0:2] = torch.tensor(struct.unpack('ii', struct.pack('Q', seed)), dtype=torch.int32)
        self._meta_host[2] = 1.0 - self.betas[0] ** self.step_count
        self._meta_host[3] = (1.0 - self.betas[1] ** self.step_count) ** 0.5

    def _graph_body(self):
        """The iteration on static buffers; everything here is captured into the CUDA graph."""
        self._meta_dev.copy_(self._meta_host, non_blocking=True)
        seed_dev = self._meta_dev[0:2].view(torch.int64)
        loss, _ = self.forward_backward(self._x_static, self._c_static, None, 0, None, seed_dev)
        dev = self._x_static.device
        with torch.cuda.device(dev):
            self._optim(self.enc, self._meta_dev[2:4])
            self._optim(self.dec, self._meta_dev[2:4])
            self.enc.module._repack(dev)
            self.dec.module._repack(dev)

    def _step_graph(self, x, c):
        key = (tuple(x.shape), float(self.loss_scale) if self.loss_scale else None)
        if self._graph is None or key != self._graph_key:
            self._x_static = torch.empty_like(x)
            self._c_static = torch.empty_like(c)
            self._graph = torch.cuda.CUDAGraph()
            self._x_static.copy_(x)
            self._c_static.copy_(c)
            self._set_meta()
            torch.cuda.synchronize()
            with torch.cuda.graph(self._graph):
                self._graph_body()
            self._graph_key = key
        self._x_static.copy_(x, non_blocking=True)
        self._c_static.copy_(c, non_blocking=True)
        self._set_meta()
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
        if self.world > 1:
            self.side.wait_stream(torch.cuda.current_stream())
            self._allreduce(self.enc, self.side)
            torch.cuda.current_stream().wait_stream(self.side)
        with torch.cuda.device(x.device):
            self._optim(self.enc)           # :330 grad_clip per network, :332 ae_opt.step()
            self._optim(self.dec)
            # refresh the tensor-core operands from the updated fp32 parameters (same allocations)
            self.enc.module._repack(x.device)
            self.dec.module._repack(x.device)
            self.enc.module._packed_key = self.dec.module._packed_key = None     # the eval handles are stale now
        self._after_step()
        return loss

    def _after_step(self):
        if self._skipped_host is not None and self._skip_event is None:
            self._skipped_host.copy_(self.skipped, non_blocking=True)
            self._skip_event = torch.cuda.Event()
            self._skip_event.record()

    def grad_norms(self):
        """(encoder, decoder) gradient L2 norms of the last step, before clipping (host floats; synchronises)."""
        return (float(self.enc.sqnorm.sqrt()) / self.world, float(self.dec.sqnorm.sqrt()) / self.world)


# ---------------------------------------------------------------------------------------------------------
# autograd wrappers: the reference's loop shape (trainer.py:246-254, 325-329) with loss.backward()
# ---------------------------------------------------------------------------------------------------------
AUTOGRAD_LOSS_SCALE = 2.0 ** 20


class _EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, noise, dropout_seed, *params):
        act, logits, _ = module.forward_train(x, noise, dropout_seed)
        ctx.module = module
        ctx.mark_non_differentiable(logits)
        return act, logits

    @staticmethod
    def backward(ctx, d_act, _d_logits):
        m = ctx.module
        names = [n for n, _ in m.named_parameters()]
        params = [p for _, p in m.named_parameters()]
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=d_act.device)
        grads, off = {}, 0
        for n, p in zip(names, params):
            grads[n] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        m.backward(d_act, grads, AUTOGRAD_LOSS_SCALE, d_act_scale=1.0)
        return (None, None, None, None) + tuple(grads[n] for n in names)


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, c, *params):
        ctx.module = module
        return module.forward_train(x, c)

    @staticmethod
    def backward(ctx, d_spec):
        m = ctx.module
        names = [n for n, _ in m.named_parameters()]
        params = [p for _, p in m.named_parameters()]
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=d_spec.device)
        grads, off = {}, 0
        for n, p in zip(names, params):
            grads[n] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        d_act = m.backward(grads, AUTOGRAD_LOSS_SCALE, d_spec=d_spec)
        return (None, d_act / AUTOGRAD_LOSS_SCALE, None) + tuple(grads[n] for n in names)


def encode_step(encoder, x, noise=None, dropout_seed=0):
    """`Trainer.encode_step` (trainer.py:246-249) with an autograd graph: returns (enc_act, enc)."""
    if noise is None:
        B, _, T = x.shape
        noise = gumbel_from_uniform(torch.rand(B, encoder.t8(T), encoder.enc_size)).to(x.device)
    return _EncodeFn.apply(encoder, x, noise, dropout_seed, *encoder.parameters())


def decode_step(decoder, enc_act, c):
    """`Trainer.decode_step` (trainer.py:251-254) with an autograd graph."""
    return _DecodeFn.apply(decoder, enc_act, c, *decoder.parameters())
