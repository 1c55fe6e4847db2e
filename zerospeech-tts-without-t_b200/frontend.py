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


def write_unit_ids(path, ids, enc_size):
    """The file `write_encodings` (convert.py:120-126) produces for one-hot units, written from the unit IDS: every line
    is enc_size ints, all '0' but a '1' at the unit's position.  One vectorised fill + one write instead of
    enc_size str() calls per unit frame (the reference formats 1024 Python ints per 100 ms of speech)."""
    ids = np.asarray(ids).reshape(-1).astype(np.int64)
    if ids.size and (ids.min() < 0 or ids.max() >= enc_size):
        raise RuntimeError(f'unit id outside [0, {enc_size})')
    line = np.frombuffer(('0 ' * enc_size)[:-1].encode() + b'\n', dtype=np.uint8)
    buf = np.tile(line, (ids.size, 1))
    buf[np.arange(ids.size), 2 * ids] = ord('1')
    with open(path, 'wb') as f:
        f.write(buf.tobytes())


def encode_to_files(path, named_specs, out_dir, ext='.txt', noise_seed=None, reference_noise_order=True):
    """`test_encode` (convert.py:342-360) over in-memory features: every (name, (L, 513) spectrogram) pair becomes
    `out_dir/name + ext` in the unit-file format the ZeroSpeech scorer reads.  All utterances go through the encoder in
    large batches and only the ids (4 bytes per unit frame) come back from the device.  `path`: an AutoencoderPath (or a
    shard.ShardedPath with gather=False semantics: each rank writes the files of its own utterances)."""
    import os
    names = [n for n, _ in named_specs]
    specs = [s for _, s in named_specs]
    enc = path.Encoder if hasattr(path, 'Encoder') else path.path.Encoder
    if enc.enc_mode != 'one_hot':
        units = path.encode_utterances(specs, reference_noise_order, noise_seed=noise_seed)
        for n, u in zip(names, units):
            write_encodings(os.path.join(out_dir, n + ext), u)
        return names
    if hasattr(path, 'shards'):          # sharded: per-rank file output, nothing is gathered
        mine, units = path.encode_utterances(specs, reference_noise_order, as_ids=True, noise_seed=noise_seed, gather=False)
    else:
        mine, units = range(len(specs)), path.encode_utterances(specs, reference_noise_order, as_ids=True, noise_seed=noise_seed)
    for i, u in zip(mine, units):
        write_unit_ids(os.path.join(out_dir, names[i] + ext), u, enc.enc_size)
    return [names[i] for i in mine]


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
      * `spec_host` may be float16: the decoder's last kernel rounds its (0, 1) output once to fp16 (<= 2.5e-4 absolute - NOT
        bit-identical, inside the path's tolerance) and half the bytes come back;
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

    def _buffers(self, T, x_dtype, layout, with_noise, spec_dtype=torch.float32):
        key = (T, x_dtype, layout, with_noise, spec_dtype)
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
                    spec=torch.empty(mb, self.dec.c_out, 8 * T8, dtype=spec_dtype, device=dev),
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
        if spec_host.dtype not in (torch.float32, torch.float16):
            raise RuntimeError('run_async: spec_host must be float32 or float16 (fp16: the output rounded once, half the download)')
        sets = self._buffers(T, x_host.dtype, layout, noise_host is not None, spec_host.dtype)
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


_COPY_POOL = None


def _copy_pool():
    """Host threads for the segment gather / result scatter of the batched front-end: a 960-segment batch moves 252 MB of
    spectrogram rows into the pinned staging buffer and as much back out, one contiguous 256 KB block per segment - numpy releases
    the GIL inside those copies, so a handful of threads multiplies the single-thread memcpy rate the host API was bound by
    (B200 box, 64 utterances x 2 000 frames through convert_utterances: 86 -> 52 ms)."""
    global _COPY_POOL
    if _COPY_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)), thread_name_prefix='zs-copy')
    return _COPY_POOL


def _parallel_ranges(pool, n, fn, min_per_task=16):
    """fn(lo, hi) over [0, n) in contiguous slices on the pool (inline when the batch is small)."""
    workers = pool._max_workers
    if n < 2 * min_per_task or workers == 1:
        fn(0, n)
        return
    tasks = min(workers, n // min_per_task)
    step = (n + tasks - 1) // tasks
    futs = [pool.submit(fn, lo, min(n, lo + step)) for lo in range(0, n, step)]
    for f in futs:
        f.result()


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
    def _plan(self, specs, want):
        """Chunk plan of every wanted utterance (convert.py:139-165): list of (utt, chunk no, T, first frame, zero-padded frames),
        plus per utterance the keep_units rule and the output frame offset of every chunk."""
        segs, keeps, out_off, out_len = [], [], [], []
        for u, spec in enumerate(specs):
            padded, plan, keep = segment_plan(len(spec), self.seg_len)
            keeps.append(keep)
            offs, o = [], 0
            for j, (s, e) in enumerate(plan):
                offs.append(o)
                o += 8 * Encoder.t8(e - s)
                if u in want:
                    segs.append((u, j, e - s, s, max(0, e - len(spec))))
            out_off.append(offs)
            out_len.append(o)
        return segs, keeps, out_off, out_len

    def _staging(self, n, T, T_out, decode):
        """Two sets of pinned host staging buffers (grow-only): segments are gathered into `x` by the host while the previous
        batch computes, results land in `spec` / `units` by asynchronous D2H and are scattered to the per-utterance arrays."""
        need_x, need_s = n * T * self.Encoder.c_in, n * T_out * self.Decoder.c_out
        st = getattr(self, '_stage', None)
        if st is None or st[0]['x'].numel() < need_x or (decode and st[0]['spec'].numel() < need_s):
            torch.cuda.synchronize(self.device)
            grow = lambda old, need: max(need, old.numel() if old is not None else 0)
            st = [dict(x=torch.empty(grow(None if st is None else st[k]['x'], need_x)).pin_memory(),
                       spec=torch.empty(grow(None if st is None else st[k]['spec'], need_s if decode else 1)).pin_memory(),
                       done=None) for k in range(2)]
            self._stage = st
        return st

    def _run(self, specs, speakers, enc_only, decode, reference_noise_order, only=None, as_ids=False, noise_seed=None,
             vocoder=None, trim=True):
        """`only`: utterance indices this process is responsible for (multi-GPU sharding, shard.py).
        `vocoder` (a dsp.GriffinLim): the decoded rows stay on the device and go straight into spectrogram2wav
        (convert.py:170) - waveforms are returned in place of spectrograms.
        Noise: `reference_noise_order` draws every chunk's Gumbel noise in the reference's call order from torch's CPU
        generator (bit-identical units; an O(all chunks) serial section on every rank); with `noise_seed` the noise is
        drawn on the device, one stream per (utterance, chunk) - independent of batching and of the sharding."""
        enc = self.Encoder
        one_hot = enc.enc_mode == 'one_hot'
        if noise_seed is not None and not one_hot:
            raise RuntimeError('device-generated noise exists for enc_mode one_hot only')
        wanted = range(len(specs)) if only is None else list(only)
        want = set(wanted)
        specs = [np.asarray(s, dtype=np.float32) for s in specs]
        segs, keeps, out_off, out_len = self._plan(specs, want)
        noises = None
        if enc.enc_mode != 'continues' and noise_seed is None and reference_noise_order:
            # the reference draws per chunk, in call order, for EVERY utterance (also those other ranks own)
            noises = {}
            for u, spec in enumerate(specs):
                for j, (s, e) in enumerate(segment_plan(len(spec), self.seg_len)[1]):
                    g = sample_gumbel(enc.noise_shape(1, e - s))
                    if u in want:
                        noises[(u, j)] = g
        c_in, c_out = enc.c_in, self.Decoder.c_out
        # per-utterance results, allocated once and filled chunk by chunk
        to_wav = decode and vocoder is not None
        res_specs = {u: np.empty((out_len[u], c_out), np.float32) for u in wanted} if decode and not to_wav else {}
        if to_wav:      # all decoded rows of the wanted utterances, utterance-major, on the device
            row0, acc = {}, 0
            for u in wanted:
                row0[u] = acc
                acc += out_len[u]
            dev_rows = torch.empty(acc, c_out, dtype=torch.float32, device=self.device)
        res_units = {u: (np.empty(out_len[u] // 8, np.int32) if one_hot else np.empty((out_len[u] // 8, enc.enc_size), np.float32))
                     for u in wanted}
        by_len = {}
        for i, sg in enumerate(segs):
            by_len.setdefault(sg[2], []).append(i)
        batches = [(T, idxs[k:k + self.max_batch]) for T, idxs in by_len.items() for k in range(0, len(idxs), self.max_batch)]
        pending = None
        if batches:
            n_max, T_max = max(len(c) for _, c in batches), max(t for t, _ in batches)
            self._staging(n_max, T_max, 8 * Encoder.t8(T_max), decode)

        pool = _copy_pool()

        def finish(job):
            chunk, T_out, T8, k, units_h = job
            self._stage[k]['done'].synchronize()
            spec_h = self._stage[k]['spec'][:len(chunk) * T_out * c_out].view(len(chunk), T_out, c_out).numpy() if decode and not to_wav else None
            u_np = units_h.numpy()

            def scatter(lo, hi):
                for n in range(lo, hi):
                    u, j = segs[chunk[n]][0], segs[chunk[n]][1]
                    o = out_off[u][j]
                    if spec_h is not None:
                        res_specs[u][o:o + T_out] = spec_h[n]
                    res_units[u][o // 8:o // 8 + T8] = u_np[n] if one_hot else u_np[n].T
            _parallel_ranges(pool, len(chunk), scatter)

        for b, (T, chunk) in enumerate(batches):
            n, T8 = len(chunk), Encoder.t8(T)
            T_out, k = 8 * T8, b % 2
            st = self._stage[k]
            if st['done'] is not None:
                st['done'].synchronize()
            xh = st['x'][:n * T * c_in].view(n, T, c_in)
            xn = xh.numpy()
            def gather(lo, hi, chunk=chunk, xn=xn, T=T):       # the chunk rows (zero-padded tail of a MIN_LEN utterance)
                for m in range(lo, hi):
                    u, _, _, s0, pad = segs[chunk[m]]
                    if pad:
                        xn[m, :T - pad] = specs[u][s0:s0 + T - pad]
                        xn[m, T - pad:] = 0.0
                    else:
                        xn[m] = specs[u][s0:s0 + T]
            _parallel_ranges(pool, n, gather)
            x = xh.to(self.device, non_blocking=True)          # (n, T, 513): consumed as is (layout 'ntc')
            noise = seeds = None
            if noises is not None:
                noise = torch.cat([noises[(segs[i][0], segs[i][1])] for i in chunk], dim=0)
            elif noise_seed is not None:                       # stream keyed by (utterance, chunk-in-utterance)
                key = torch.tensor([(segs[i][0] << 20) + segs[i][1] for i in chunk], dtype=torch.int64)
                seeds = (key.to(self.device) * 0x2545F4914F6CDD1D + int(noise_seed))
            elif enc.enc_mode != 'continues':
                noise = sample_gumbel(enc.noise_shape(n, T))
            act, _, ids = enc.encode(x, noise, layout='ntc', noise_seeds=seeds, want_act=not one_hot, want_logits=not one_hot)
            if decode:
                c = torch.tensor([speakers[segs[i][0]] for i in chunk], dtype=torch.int64).to(self.device)
                x_dec = self._decode(act, ids, c, enc_only)
                if to_wav:
                    dst = torch.tensor([row0[segs[i][0]] + out_off[segs[i][0]][segs[i][1]] for i in chunk], dtype=torch.int64)
                    rows = (dst.to(self.device)[:, None] + torch.arange(T_out, device=self.device)[None, :]).reshape(-1)
                    dev_rows.index_copy_(0, rows, x_dec.permute(0, 2, 1).reshape(-1, c_out))
                else:
                    st['spec'][:n * T_out * c_out].view(n, T_out, c_out).copy_(x_dec.permute(0, 2, 1), non_blocking=True)
            # one_hot: 4 bytes per unit frame leave the device instead of 4 * enc_size
            units_h = torch.empty((n, T8) if one_hot else (n, enc.enc_size, T8), dtype=torch.int32 if one_hot else torch.float32).pin_memory()
            units_h.copy_(ids if one_hot else act, non_blocking=True)
            st['done'] = torch.cuda.Event()
            st['done'].record()
            if pending is not None:
                finish(pending)                                # scatter the previous batch while this one runs
            pending = (chunk, T_out, T8, k, units_h)
        if pending is not None:
            finish(pending)
        check_range(self.device)
        out_units, out_specs = [], []
        for u in wanted:
            e = res_units[u]
            if keeps[u] is not None:
                e = e[:keeps[u]]
            if one_hot and not as_ids:
                e = one_hot_rows(e, enc.enc_size)
            out_units.append(e)
            if decode and not to_wav:
                out_specs.append(res_specs[u])
        if to_wav:
            out_specs = vocoder.synthesize(dev_rows, [out_len[u] for u in wanted], trim=trim)
        return out_specs, out_units

    def encode_utterances(self, specs, reference_noise_order=True, only=None, as_ids=False, noise_seed=None):
        """encode() for a list of (L, 513) spectrograms -> list of (n_units, enc_size) arrays (convert.py:183-221), or of
        (n_units,) int32 id arrays with `as_ids` (one_hot)."""
        return self._run(specs, None, True, False, reference_noise_order, only, as_ids, noise_seed)[1]

    def convert_utterances(self, specs, target_speakers, enc_only=True, reference_noise_order=True, only=None, as_ids=False,
                           noise_seed=None, vocoder=None, trim=True):
        """convert() (convert.py:128-180): -> (list of (L', 513) spectrograms, list of unit arrays); with `vocoder`
        (a `dsp.GriffinLim`) the first list holds the waveforms of spectrogram2wav (convert.py:170) instead - the decoded
        spectrograms then never leave the device."""
        return self._run(specs, list(target_speakers), enc_only, True, reference_noise_order, only, as_ids, noise_seed, vocoder, trim)
