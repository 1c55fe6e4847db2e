"""Batched front-end for the autoencoder path: the reference's inference drivers
(convert.py:70-83 convert_x/encode_x, :128-221 convert()/encode() chunking,
trainer.py:180-228 set_eval/test_step/encoder_test_step) restated so that all
segments of all utterances go through the CUDA path in large batches instead of
one batch-1 call (+ one device->host sync) per 128-frame chunk.

Segments are independent in the reference (InstanceNorm is per sample, the GRU
state is re-zeroed per call), so batching them is exact.
"""
import numpy as np
import torch

from .model import Decoder, Encoder, check_range, sample_gumbel

MIN_LEN = 9  # convert.py:36


def segment_plan(n_frames, seg_len):
    """Chunking rule of convert()/encode() (convert.py:139-165, 189-213).

    Returns (padded_len, [(start, stop), ...], keep_units): the utterance is zero-padded to
    `padded_len` (>= MIN_LEN) frames, the model sees one segment per (start, stop) slice, and
    only the first `keep_units` unit frames are kept when the utterance had to be padded.
    Quirks kept: the tail segment drops the utterance's last frame (`spec[idx:-1]`) and is
    seg_len..2*seg_len-1 frames long; shorter leftovers are discarded."""
    padded = max(n_frames, MIN_LEN)
    keep = MIN_LEN // 8 if n_frames < MIN_LEN else None
    if padded <= seg_len:
        return padded, [(0, padded)], keep
    plan = []
    for idx in range(0, padded, seg_len):
        stop = padded - 1 if idx + 2 * seg_len > padded else idx + seg_len
        if stop - idx >= seg_len:
            plan.append((idx, stop))
        elif idx == 0:
            raise RuntimeError('Please check if input is too short!')
    return padded, plan, None


def write_encodings(path, encodings):
    """convert.py:120-126: one line per unit frame, enc_size space-separated ints."""
    enc = np.asarray(encodings)
    with open(path, 'w') as f:
        for row in enc:
            f.write(' '.join(str(int(e)) for e in row) + '\n')


class StreamingResynthesizer:
    """Host-to-host encode -> decode over many 128-frame segments with copies overlapped with compute.

    The per-chunk reference loop pays one H2D, one launch train and one D2H *sync* per segment
    (convert.py:70-76, trainer.py:221).  Here segments go through in micro-batches on three CUDA streams:
    H2D of micro-batch i+1 and D2H of micro-batch i-1 run on the two copy engines while micro-batch i computes;
    `n_buffers` device buffer sets bound the micro-batches in flight.  Host tensors must be pinned for the copies
    to be asynchronous.  Micro-batches of 960 segments of 128 frames tile the B200 (25.95 waves of 148 CTAs on the
    wide layers, exactly 2 waves of 15 sixty-four-sequence GRU clusters): such calls run at ~15 M frames/s against
    12 M at 222 segments (4.4 MB of workspace per segment).

    Byte-saving inputs, all bit-identical on the spectrogram side:
      * `x_host` may be float16 (the path rounds its input to fp16 operands first thing: same results, half the upload);
      * `layout='ntc'`: x_host is (S, T, c_in) - the layout of the HDF5 features and of Trainer.test_step's argument
        (trainer.py:196 permutes it) - and is consumed without a transpose copy;
      * `noise_host=None` with `noise_seed=<int>`: the Gumbel noise of the bottleneck is drawn on the device, one
        counter-based stream per segment (seed + global segment number) - same distribution as model/model.py:95-98 but
        not the reference's CPU-generator stream (the throughput mode; pass `noise_host` for reference-exact units)."""

    def __init__(self, encoder: Encoder, decoder: Decoder, micro_batch=960, n_buffers=4, device='cuda'):
        self.enc, self.dec = encoder, decoder
        self.mb, self.nbuf = micro_batch, n_buffers
        self.device = torch.device(device)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self._bufs = None
        self._issued = 0          # micro-batches issued over the life of the object (buffer sets rotate across calls)

    def _buffers(self, T, x_dtype, layout, with_noise):
        key = (T, x_dtype, layout, with_noise)
        if self._bufs is None or self._bufs[0] != key:
            for st in (self.s_in, self.s_cmp, self.s_out):    # earlier calls may still be using the old buffers
                st.synchronize()
            mb, dev, enc = self.mb, self.device, self.enc
            T8 = Encoder.t8(T)
            xshape = (mb, enc.c_in, T) if layout == 'nct' else (mb, T, enc.c_in)
            sets = []
            for _ in range(self.nbuf):
                sets.append(dict(
                    x=torch.empty(xshape, dtype=x_dtype, device=dev), c=torch.empty(mb, dtype=torch.int64, device=dev),
                    noise=torch.empty(enc.noise_shape(mb, T), device=dev) if with_noise else None,
                    seeds=None if with_noise else torch.empty(mb, dtype=torch.int64, device=dev),
                    spec=torch.empty(mb, self.dec.c_out, 8 * T8, device=dev),
                    ids=torch.empty(mb, T8, dtype=torch.int32, device=dev), used=False,
                    loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event()))
            self._bufs = (key, sets)
            self._arange = torch.arange(mb, dtype=torch.int64, device=dev)
        return self._bufs[1]

    @torch.no_grad()
    def run_async(self, x_host, c_host, spec_host, ids_host=None, noise_host=None, layout='nct', noise_seed=None, segment0=0):
        """x_host (S, c_in, T) [layout 'nct'] or (S, T, c_in) ['ntc'], fp32 or fp16; c_host (S,) int64; noise_host
        (S, T8, enc_size) or None + noise_seed -> spec_host (S, c_out, T'), ids_host (S, T8) int32.  `segment0` = global
        number of the call's first segment (device-noise streams are keyed by it, so results do not depend on how a
        job is cut into calls).  Returns a CUDA event that completes when everything has landed in the host tensors;
        nothing here waits on the host, and consecutive calls overlap (the upload of call k+1 runs under the compute
        and download of call k), so the host tensors of a call must stay untouched until its event has completed."""
        S = x_host.shape[0]
        T = x_host.shape[2] if layout == 'nct' else x_host.shape[1]
        if noise_host is None and noise_seed is None and self.enc.enc_mode != 'continues':
            raise RuntimeError('run_async: pass noise_host (reference-exact draws) or noise_seed (device-generated noise)')
        if x_host.dtype not in (torch.float32, torch.float16):
            raise RuntimeError('run_async: x_host must be float32 or float16')
        sets = self._buffers(T, x_host.dtype, layout, noise_host is not None)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))      # whatever prepared the inputs on the caller's stream
        self.s_in.wait_event(ready)
        n_mb = (S + self.mb - 1) // self.mb
        for i in range(n_mb):
            s0, s1 = i * self.mb, min(S, (i + 1) * self.mb)
            n = s1 - s0
            b = sets[self._issued % self.nbuf]
            self._issued += 1
            with torch.cuda.stream(self.s_in):
                if b['used']:
                    self.s_in.wait_event(b['computed'])       # the compute that last read this input set is done
                b['x'][:n].copy_(x_host[s0:s1], non_blocking=True)
                b['c'][:n].copy_(c_host[s0:s1], non_blocking=True)
                if noise_host is not None:
                    b['noise'][:n].copy_(noise_host[s0:s1], non_blocking=True)
                b['loaded'].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(b['loaded'])
                if b['used']:
                    self.s_cmp.wait_event(b['drained'])       # its previous outputs have left the device
                if noise_host is not None:
                    noise, seeds = b['noise'][:n], None
                else:                                          # splitmix-style spread of (seed, global segment number)
                    noise = None
                    seeds = torch.add(self._arange[:n], segment0 + s0, out=b['seeds'][:n]).mul_(0x2545F4914F6CDD1D).add_(int(noise_seed))
                _, _, ids = self.enc.encode(b['x'][:n], noise, layout=layout, noise_seeds=seeds, want_act=False, want_logits=False)
                self.dec.decode(None, b['c'][:n], unit_ids=ids, out=b['spec'][:n])
                b['ids'][:n].copy_(ids)
                b['computed'].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(b['computed'])
                spec_host[s0:s1].copy_(b['spec'][:n], non_blocking=True)
                if ids_host is not None:
                    ids_host[s0:s1].copy_(b['ids'][:n], non_blocking=True)
                b['drained'].record(self.s_out)
            b['used'] = True
        done = torch.cuda.Event()
        done.record(self.s_out)          # downloads are issued in order on one stream: the last one closes the call
        return done

    def run(self, x_host, c_host, spec_host, ids_host=None, noise_host=None, **kw):
        """`run_async` + the caller's stream waits for the results (stream-ordered, like a torch op)."""
        done = self.run_async(x_host, c_host, spec_host, ids_host, noise_host, **kw)
        torch.cuda.current_stream(self.device).wait_event(done)
        return done

    def check_range(self):
        """Waits for everything issued and raises `OperandRangeError` if an fp16 operand saturated (see model.check_range)."""
        for st in (self.s_in, self.s_cmp, self.s_out):
            st.synchronize()
        return check_range(self.device)


def one_hot_rows(ids, enc_size):
    """(n,) unit ids -> the (n, enc_size) 0/1 float32 rows the reference's encode() returns (convert.py:183-221)."""
    ids = np.asarray(ids).reshape(-1)
    out = np.zeros((ids.shape[0], enc_size), np.float32)
    out[np.arange(ids.shape[0]), ids] = 1.0
    return out


class AutoencoderPath:
    """The Trainer's inference surface for this path, backed by the B200 modules.

    Mirrors Trainer.encoder_test_step / Trainer.test_step (trainer.py:194-228) and adds the batched
    `encode_utterances` / `convert_utterances` that replace the per-chunk Python loops of convert.py.
    `generator` is the TTS patcher of trainer.py:70-81: a second `Decoder` for g_mode naive / targeted /
    targeted_residual, a `patchers.Enhanced_Generator` / `patchers.Spectrogram_Patcher` for enhanced / spectrogram."""

    def __init__(self, encoder: Encoder, decoder: Decoder, generator=None, g_mode='targeted',
                 n_speakers=102, n_target_speakers=2, seg_len=128, max_batch=960, device='cuda'):
        self.Encoder, self.Decoder, self.Generator = encoder, decoder, generator
        self.g_mode, self.seg_len, self.max_batch = g_mode, seg_len, max_batch
        self.shift = n_speakers - n_target_speakers        # trainer.py:181 testing_shift_c
        self.n_target_speakers = n_target_speakers
        self.device = torch.device(device)
        for m in (encoder, decoder, generator):
            if m is not None:
                m.to(self.device).eval()

    # ---- trainer.py:194-228, same signatures and return types ------------------------------
    def test_step(self, x, c, enc_only=False, noise=None):
        """x: (B, T, 513) float tensor, c: (B,) speaker ids -> (x_dec (B, 513, T') numpy, enc (B, enc, T8) numpy)."""
        x = x.to(self.device, torch.float32)
        c = torch.as_tensor(c).to(self.device).view(-1)
        enc, _, ids = self.Encoder.encode(x, noise, layout='ntc')       # trainer.py:196 permutes; the kernel reads (B, T, C)
        x_dec = self._decode(enc, ids, c, enc_only)
        out = x_dec.cpu().numpy(), enc.cpu().numpy()
        check_range(self.device)
        return out

    def encoder_test_step(self, x, noise=None):
        x = x.to(self.device, torch.float32)
        enc, _, _ = self.Encoder.encode(x, noise, layout='ntc')
        out = enc.cpu().numpy()
        check_range(self.device)
        return out

    def _decode(self, enc, ids, c, enc_only, check_targets=True):
        use_ids = ids is not None
        x_dec = self.Decoder.decode(None if use_ids else enc, c, unit_ids=ids if use_ids else None)
        if enc_only:
            return x_dec
        if self.Generator is None:
            raise RuntimeError('enc_only=False needs a Generator (trainer.py:200-217)')
        if self.g_mode == 'naive':                               # trainer.py:206-207
            self.Generator.decode(None if use_ids else enc, c, unit_ids=ids if use_ids else None, out=x_dec, accumulate=1)
            return x_dec
        cg = c - self.shift
        if check_targets and (int(cg.min()) < 0 or int(cg.max()) >= self.n_target_speakers):
            raise RuntimeError('This generator can only convert to target speakers!')   # trainer.py:202-203
        if self.g_mode in ('targeted', 'targeted_residual'):     # :208-211
            self.Generator.decode(None if use_ids else enc, cg, unit_ids=ids if use_ids else None, out=x_dec,
                                  accumulate=1 if self.g_mode == 'targeted' else 2)
        elif self.g_mode in ('enhanced', 'spectrogram'):         # :212-213  x_dec += Generator(x_dec, c - shift)
            self.Generator.patch(x_dec, cg, out=x_dec, accumulate=1)
        else:
            raise NotImplementedError(f'Invalid generator mode {self.g_mode!r}')
        return x_dec

    # ---- convert.py:128-221 batched -----------------------------------------------------------
    def _segments(self, specs):
        segs = []          # (utt index, order in utt, T, padded spec slice)
        keeps = []
        for u, spec in enumerate(specs):
            spec = np.asarray(spec, dtype=np.float32)
            padded, plan, keep = segment_plan(len(spec), self.seg_len)
            if padded > len(spec):
                spec = np.concatenate([spec, np.zeros((padded - len(spec), spec.shape[1]), np.float32)], axis=0)
            keeps.append(keep)
            for j, (s, e) in enumerate(plan):
                segs.append((u, j, e - s, spec[s:e]))
        return segs, keeps

    def _run(self, specs, speakers, enc_only, decode, reference_noise_order, only=None, as_ids=False, noise_seed=None):
        """`only`: utterance indices this process is responsible for (multi-GPU sharding, shard.py).
        Noise: `reference_noise_order` draws every chunk's Gumbel noise in the reference's call order from torch's CPU
        generator (bit-identical units; an O(all chunks) serial section on every rank); with `noise_seed` the noise is
        drawn on the device, one stream per (utterance, chunk) - independent of batching and of the sharding."""
        segs, keeps = self._segments(specs)
        enc = self.Encoder
        one_hot = enc.enc_mode == 'one_hot'
        noises = None
        if enc.enc_mode != 'continues' and noise_seed is None and reference_noise_order:
            noises = [sample_gumbel(enc.noise_shape(1, T)) for (_, _, T, _) in segs]
        if noise_seed is not None and not one_hot:
            raise RuntimeError('device-generated noise exists for enc_mode one_hot only')
        wanted = range(len(specs)) if only is None else list(only)
        want = set(wanted)
        by_len = {}
        for i, (u, _, T, _) in enumerate(segs):
            if u in want:
                by_len.setdefault(T, []).append(i)
        units = [None] * len(segs)
        outs = [None] * len(segs)
        for T, idxs in by_len.items():
            for k in range(0, len(idxs), self.max_batch):
                chunk = idxs[k:k + self.max_batch]
                x = torch.from_numpy(np.stack([segs[i][3] for i in chunk])).pin_memory()     # (n, T, 513): consumed as is
                x = x.to(self.device, non_blocking=True)
                noise = seeds = None
                if noises is not None:
                    noise = torch.cat([noises[i] for i in chunk], dim=0)
                elif noise_seed is not None:       # stream keyed by (utterance, chunk-in-utterance)
                    key = torch.tensor([(segs[i][0] << 20) + segs[i][1] for i in chunk], dtype=torch.int64)
                    seeds = (key.to(self.device) * 0x2545F4914F6CDD1D + int(noise_seed))
                elif enc.enc_mode != 'continues':
                    noise = sample_gumbel(enc.noise_shape(len(chunk), T))
                act, _, ids = enc.encode(x, noise, layout='ntc', noise_seeds=seeds, want_act=not one_hot, want_logits=not one_hot)
                if decode:
                    c = torch.tensor([speakers[segs[i][0]] for i in chunk], dtype=torch.int64, device=self.device)
                    x_dec = self._decode(act, ids, c, enc_only).permute(0, 2, 1).cpu().numpy()
                # one_hot: 4 bytes per unit frame leave the device instead of 4 * enc_size
                u_np = ids.cpu().numpy() if one_hot else act.permute(0, 2, 1).cpu().numpy()
                for n, i in enumerate(chunk):
                    units[i] = u_np[n]
                    if decode:
                        outs[i] = x_dec[n]
        check_range(self.device)
        res_units, res_specs = [], []
        for u in wanted:
            mine = sorted((j, i) for i, (uu, j, _, _) in enumerate(segs) if uu == u)
            e = np.concatenate([units[i] for _, i in mine], axis=0)
            if keeps[u] is not None:
                e = e[:keeps[u]]
            if one_hot and not as_ids:
                e = one_hot_rows(e, enc.enc_size)
            res_units.append(e)
            if decode:
                res_specs.append(np.concatenate([outs[i] for _, i in mine], axis=0))
        return res_specs, res_units

    def encode_utterances(self, specs, reference_noise_order=True, only=None, as_ids=False, noise_seed=None):
        """encode() for a list of (L, 513) spectrograms -> list of (n_units, enc_size) arrays (convert.py:183-221), or of
        (n_units,) int32 id arrays with `as_ids` (one_hot)."""
        return self._run(specs, None, True, False, reference_noise_order, only, as_ids, noise_seed)[1]

    def convert_utterances(self, specs, target_speakers, enc_only=True, reference_noise_order=True, only=None, as_ids=False,
                           noise_seed=None):
        """convert() up to (not including) Griffin-Lim: -> (list of (L', 513) spectrograms, list of unit arrays)."""
        return self._run(specs, list(target_speakers), enc_only, True, reference_noise_order, only, as_ids, noise_seed)
