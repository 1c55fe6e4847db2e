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

from .model import Decoder, Encoder, sample_gumbel

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
    H2D of micro-batch i+1 and D2H of micro-batch i-1 run on the two copy engines while micro-batch i computes.
    Host tensors must be pinned for the copies to be asynchronous.  Micro-batches of 960 segments of 128 frames tile the
    B200 (25.95 waves of 148 CTAs on the wide layers, exactly 2 waves of 15 sixty-four-sequence GRU clusters): such calls
    run at ~15 M frames/s against 12 M at 222 segments (4.4 MB of workspace per segment)."""

    def __init__(self, encoder: Encoder, decoder: Decoder, micro_batch=960, n_buffers=4, device='cuda'):
        self.enc, self.dec = encoder, decoder
        self.mb, self.nbuf = micro_batch, n_buffers
        self.device = torch.device(device)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self._bufs = None
        self._issued = 0          # micro-batches issued over the life of the object (buffer sets rotate across calls)

    def _buffers(self, T, with_noise):
        key = (T, with_noise)
        if self._bufs is None or self._bufs[0] != key:
            for st in (self.s_in, self.s_cmp, self.s_out):    # earlier calls may still be using the old buffers
                st.synchronize()
            mb, dev, enc = self.mb, self.device, self.enc
            T8 = Encoder.t8(T)
            sets = []
            for _ in range(self.nbuf):
                sets.append(dict(
                    x=torch.empty(mb, enc.c_in, T, device=dev), c=torch.empty(mb, dtype=torch.int64, device=dev),
                    noise=torch.empty(enc.noise_shape(mb, T), device=dev) if with_noise else None,
                    spec=torch.empty(mb, self.dec.c_out, 8 * T8, device=dev),
                    ids=torch.empty(mb, T8, dtype=torch.int32, device=dev), used=False,
                    loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event()))
            self._bufs = (key, sets)
        return self._bufs[1]

    @torch.no_grad()
    def run_async(self, x_host, c_host, spec_host, ids_host=None, noise_host=None):
        """x_host (S, c_in, T) fp32, c_host (S,) int64, optional noise_host (S, T8, enc_size) -> spec_host (S, c_out, T'),
        ids_host (S, T8) int32.  Returns a CUDA event that completes when everything has landed in the host tensors;
        nothing here waits on the host, and consecutive calls overlap (the upload of call k+1 runs under the compute
        and download of call k), so the host tensors of a call must stay untouched until its event has completed."""
        S, _, T = x_host.shape
        sets = self._buffers(T, noise_host is not None)
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
                noise = b['noise'][:n] if noise_host is not None else None
                _, _, ids = self.enc.encode(b['x'][:n], noise)
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

    def run(self, x_host, c_host, spec_host, ids_host=None, noise_host=None):
        """`run_async` + the caller's stream waits for the results (stream-ordered, like a torch op)."""
        done = self.run_async(x_host, c_host, spec_host, ids_host, noise_host)
        torch.cuda.current_stream(self.device).wait_event(done)
        return done


class AutoencoderPath:
    """The Trainer's inference surface for this path, backed by the B200 modules.

    Mirrors Trainer.encoder_test_step / Trainer.test_step (trainer.py:194-228) and adds the batched
    `encode_utterances` / `convert_utterances` that replace the per-chunk Python loops of convert.py."""

    def __init__(self, encoder: Encoder, decoder: Decoder, generator: Decoder = None, g_mode='targeted',
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
        x = x.to(self.device, torch.float32).permute(0, 2, 1)
        c = torch.as_tensor(c).to(self.device).view(-1)
        enc, _, ids = self.Encoder.encode(x, noise)
        x_dec = self._decode(enc, ids, c, enc_only)
        return x_dec.cpu().numpy(), enc.cpu().numpy()

    def encoder_test_step(self, x, noise=None):
        x = x.to(self.device, torch.float32).permute(0, 2, 1)
        enc, _ = self.Encoder(x, noise)
        return enc.cpu().numpy()

    def _decode(self, enc, ids, c, enc_only, check_targets=True):
        use_ids = ids is not None
        x_dec = self.Decoder.decode(None if use_ids else enc, c, unit_ids=ids if use_ids else None)
        if enc_only:
            return x_dec
        if self.Generator is None:
            raise RuntimeError('enc_only=False needs a Generator (trainer.py:200-217)')
        if self.g_mode == 'naive':
            cg, acc = c, 1
        elif self.g_mode in ('targeted', 'targeted_residual'):
            cg = c - self.shift
            if check_targets and (int(cg.min()) < 0 or int(cg.max()) >= self.n_target_speakers):
                raise RuntimeError('This generator can only convert to target speakers!')   # trainer.py:202-203
            acc = 1 if self.g_mode == 'targeted' else 2
        else:
            raise NotImplementedError(f'g_mode {self.g_mode!r} is outside the autoencoder hot path')
        self.Generator.decode(None if use_ids else enc, cg, unit_ids=ids if use_ids else None, out=x_dec,
                              accumulate=acc)
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

    def _run(self, specs, speakers, enc_only, decode, reference_noise_order, only=None):
        """`only`: utterance indices this process is responsible for (multi-GPU sharding, shard.py); the noise of
        every chunk is still drawn, in order, so the draws match a single-process run."""
        segs, keeps = self._segments(specs)
        enc_size = self.Encoder.enc_size
        # the reference draws the Gumbel noise per chunk, in call order, from the CPU generator
        noises = None
        if reference_noise_order and self.Encoder.enc_mode != 'continues':
            noises = [sample_gumbel(self.Encoder.noise_shape(1, T)) for (_, _, T, _) in segs]
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
                x = torch.from_numpy(np.stack([segs[i][3] for i in chunk])).pin_memory()
                x = x.to(self.device, non_blocking=True).permute(0, 2, 1).contiguous()
                if noises is not None:
                    noise = torch.cat([noises[i] for i in chunk], dim=0)
                elif self.Encoder.enc_mode == 'continues':
                    noise = None
                else:
                    noise = sample_gumbel(self.Encoder.noise_shape(len(chunk), T))
                enc, _, ids = self.Encoder.encode(x, noise)
                if decode:
                    c = torch.tensor([speakers[segs[i][0]] for i in chunk], dtype=torch.int64, device=self.device)
                    x_dec = self._decode(enc, ids, c, enc_only).permute(0, 2, 1).cpu().numpy()
                enc_np = enc.permute(0, 2, 1).cpu().numpy()
                for n, i in enumerate(chunk):
                    units[i] = enc_np[n]
                    if decode:
                        outs[i] = x_dec[n]
        res_units, res_specs = [], []
        for u in wanted:
            mine = sorted((j, i) for i, (uu, j, _, _) in enumerate(segs) if uu == u)
            e = np.concatenate([units[i] for _, i in mine], axis=0)
            if keeps[u] is not None:
                e = e[:keeps[u]]
            res_units.append(e)
            if decode:
                res_specs.append(np.concatenate([outs[i] for _, i in mine], axis=0))
        return res_specs, res_units

    def encode_utterances(self, specs, reference_noise_order=True, only=None):
        """encode() for a list of (L, 513) spectrograms -> list of (n_units, enc_size) arrays."""
        return self._run(specs, None, True, False, reference_noise_order, only)[1]

    def convert_utterances(self, specs, target_speakers, enc_only=True, reference_noise_order=True, only=None):
        """convert() up to (not including) Griffin-Lim: -> (list of (L', 513) spectrograms, list of unit arrays)."""
        return self._run(specs, list(target_speakers), enc_only, True, reference_noise_order, only)
