"""The DSP steps either side of the autoencoder path, on the GPU:

  * `spectrogram2wav` / `GriffinLim`   convert.py:39-62  - de-normalise, Griffin-Lim x n_iter, de-pre-emphasis, trim
  * `get_spectrograms`                 preprocess.py:231-256 - pre-emphasis, STFT, log-magnitude, normalise

with the reference's constants (hps/hps.py:22-33).  Both run hand-written sm_100a kernels of libzsae.so
(csrc/stft.cuh: a shared-memory radix-8 FFT, one fused kernel per Griffin-Lim iteration); spectrograms coming from the
Decoder never leave the device on their way to the vocoder.  librosa (which the reference calls for stft / istft / trim)
is not part of the reference tree: its published algorithm (centred frames, reflect padding, window sum-of-squares
normalisation) is what these kernels implement and what the tests' CPU restatement checks them against; file decoding
and resampling (`librosa.load`) stay with the caller.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import _ptr, _stream

SR, N_FFT, HOP, WIN, N_BIN = 16000, 1024, 200, 800, 513          # hps/hps.py:22-28
N_ITER, PREEMPHASIS, MAX_DB, REF_DB = 300, 0.97, 100, 20         # hps/hps.py:30-33
TRIM_TOP_DB, TRIM_FRAME, TRIM_HOP = 60, 2048, 512                # librosa.effects.trim defaults (convert.py:61)


def _prefix(counts, dev):
    out = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=out[1:])
    if out[-1] >= 2 ** 31:
        raise RuntimeError('batch too large for 32-bit offsets: split it')
    return out


class GriffinLim:
    """spectrogram2wav for batches of utterances.  Keeps one workspace allocation per device across calls."""

    def __init__(self, device='cuda', n_iter=N_ITER, preemphasis=PREEMPHASIS):
        self.device = torch.device(device)
        self.n_iter, self.preemphasis = n_iter, preemphasis
        self._ws = None
        self.tile = None

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    @torch.no_grad()
    def synthesize(self, spec, n_frames, trim=True, to_host=True):
        """spec: (sum(n_frames), 513) float32 CUDA tensor of normalised log-magnitude rows, utterances back to back
        (the Decoder's (B, 513, T) output transposed to frames-major); n_frames: frames per utterance (each >= 4).
        Returns the list of waveforms (float32 numpy arrays when `to_host`, else device tensors), trimmed like
        librosa.effects.trim(top_db=60) when `trim`."""
        lib = _lib.lib()
        n_frames = [int(n) for n in n_frames]
        if min(n_frames) < 4:
            raise RuntimeError('GriffinLim: every utterance needs at least 4 frames (reflect padding of the 1024-sample frames)')
        dev = self.device
        spec = spec.to(dev, torch.float32).contiguous()
        if spec.dim() != 2 or spec.shape[1] != N_BIN or spec.shape[0] != sum(n_frames):
            raise RuntimeError(f'GriffinLim: expected ({sum(n_frames)}, {N_BIN}) rows, got {tuple(spec.shape)}')
        with torch.cuda.device(dev):
            if self.tile is None:
                self.tile = lib.zs_stft_tile_frames()
            U = len(n_frames)
            samples = [HOP * (n - 1) for n in n_frames]
            tiles = [(n + self.tile - 1) // self.tile for n in n_frames]
            pf = [1 + s // TRIM_HOP for s in samples]
            fs, ss, ts, ps = _prefix(n_frames, dev), _prefix(samples, dev), _prefix(tiles, dev), _prefix(pf, dev)
            meta = torch.from_numpy(np.concatenate([fs, ss, ts, ps]).astype(np.int32)).to(dev)
            total_frames, total_samples, total_tiles = int(fs[-1]), int(ss[-1]), int(ts[-1])
            wav = torch.empty(total_samples, dtype=torch.float32, device=dev)
            ws = self._workspace(lib.zs_griffin_lim_workspace_bytes(total_frames, total_samples))
            _lib.check(lib.zs_griffin_lim(_ptr(spec), _ptr(meta), U, total_frames, total_samples, total_tiles, int(self.n_iter),
                                          float(self.preemphasis), _ptr(wav), _ptr(ws), ws.numel(), _stream()))
            bounds = [(0, s) for s in samples]
            if trim:
                power = torch.empty(int(ps[-1]), dtype=torch.float32, device=dev)
                sample_start = C.c_void_p(meta.data_ptr() + 4 * (U + 1))
                pframe_start = C.c_void_p(meta.data_ptr() + 4 * 3 * (U + 1))
                _lib.check(lib.zs_frame_power(_ptr(wav), sample_start, pframe_start, U, max(pf), _ptr(power), _stream()))
                bounds = trim_bounds_from_power(power.cpu().numpy(), ps, samples)
        out = []
        host = wav.cpu().numpy() if to_host else None
        for u, (a, b) in enumerate(bounds):
            o = int(ss[u])
            out.append(host[o + a:o + b].copy() if to_host else wav[o + a:o + b])
        return out


def trim_bounds_from_power(power, pframe_start, samples, top_db=TRIM_TOP_DB):
    """librosa.effects.trim's index logic on the per-frame mean-square power the device computed:
    power_to_db(mse, ref=max, amin=1e-10) > -top_db; keep [first * hop, min(len, (last + 1) * hop))."""
    bounds = []
    for u, L in enumerate(samples):
        mse = power[int(pframe_start[u]):int(pframe_start[u + 1])].astype(np.float64)
        db = 10.0 * np.log10(np.maximum(1e-10, mse)) - 10.0 * np.log10(max(1e-10, float(mse.max()) if mse.size else 0.0))
        nz = np.flatnonzero(db > -top_db)
        bounds.append((0, 0) if nz.size == 0 else (int(nz[0] * TRIM_HOP), int(min(L, (nz[-1] + 1) * TRIM_HOP))))
    return bounds


def spectrogram2wav(mag, n_iter=N_ITER, trim=True, device='cuda'):
    """convert.py:55-62 for ONE utterance: mag (T, 513) normalised spectrogram (numpy or tensor) -> float32 numpy waveform."""
    mag = torch.as_tensor(np.asarray(mag, np.float32) if not torch.is_tensor(mag) else mag)
    return GriffinLim(device, n_iter).synthesize(mag.to(device), [mag.shape[0]], trim=trim)[0]


@torch.no_grad()
def get_spectrograms(wavs, device='cuda', dtype=torch.float32, to_host=True, preemphasis=PREEMPHASIS):
    """preprocess.py:231-256 for a list of 16 kHz float waveforms (already loaded and trimmed: `librosa.load` /
    `effects.trim` stay with the caller) -> list of (T, 513) normalised log-magnitude spectrograms, T = 1 + len // 200.
    dtype float16 gives the rows in the encoder's byte-saving input format (`Encoder.encode(..., layout='ntc')`)."""
    lib = _lib.lib()
    dev = torch.device(device)
    lens = [int(len(w)) for w in wavs]
    if min(lens) < N_FFT // 2 + 1:
        raise RuntimeError(f'get_spectrograms: a waveform shorter than {N_FFT // 2 + 1} samples cannot be reflect-padded')
    if dtype not in (torch.float32, torch.float16):
        raise RuntimeError('get_spectrograms: dtype must be float32 or float16')
    with torch.cuda.device(dev):
        tile = lib.zs_stft_tile_frames()
        frames = [1 + n // HOP for n in lens]
        tiles = [(n + tile - 1) // tile for n in frames]
        fs, ss, ts = _prefix(frames, dev), _prefix(lens, dev), _prefix(tiles, dev)
        meta = torch.from_numpy(np.concatenate([fs, ss, ts]).astype(np.int32)).to(dev)
        flat = torch.from_numpy(np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in wavs])).to(dev) \
            if not torch.is_tensor(wavs[0]) else torch.cat([w.to(dev, torch.float32).reshape(-1) for w in wavs])
        spec = torch.empty(int(fs[-1]), N_BIN, dtype=dtype, device=dev)
        s32, s16 = (spec, None) if dtype == torch.float32 else (None, spec)
        _lib.check(lib.zs_spectrogram(_ptr(flat), _ptr(meta), len(wavs), int(ts[-1]), float(preemphasis), _ptr(s32), _ptr(s16), _stream()))
    parts = [spec[int(fs[u]):int(fs[u + 1])] for u in range(len(wavs))]
    return [p.cpu().numpy() for p in parts] if to_host else parts
