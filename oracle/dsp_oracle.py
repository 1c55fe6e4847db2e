"""ORACLE - test infrastructure, NOT product code.

numpy restatement of the two DSP steps either side of the autoencoder path:

  * `get_spectrograms`  preprocess.py:227-258  (pre-emphasis, STFT 1024 / hop 200 / Hann 800, |.|, dB, normalise)
  * `spectrogram2wav`   convert.py:39-62       (de-normalise, Griffin-Lim x 300, de-pre-emphasis, trim)
    with constants hps/hps.py:17-36.

Both call `librosa` (stft / istft / effects.trim), which is NOT installed in this image and not vendored by the reference
(the reference pins no version; its README targets 2019-era librosa 0.6/0.7).  Parity is therefore anchored on librosa's
published algorithm, restated here function by function:

  librosa.stft(y, n_fft, hop_length, win_length, window='hann', center=True, pad_mode='reflect')
      window = scipy.signal.get_window('hann', win_length, fftbins=True), zero-padded (centred) to n_fft;
      y reflect-padded by n_fft // 2 on both sides; frame j = y_pad[j * hop : j * hop + n_fft] * window;
      rfft -> (1 + n_fft // 2, 1 + len(y) // hop) complex64.
  librosa.istft(S, hop_length, win_length, window='hann', center=True, length=None)
      per frame: irfft(S[:, i]) * padded window, overlap-added at i * hop into n_fft + hop * (n_frames - 1) samples;
      divided by the window sum-of-squares where that exceeds tiny(float32); n_fft // 2 samples trimmed from both ends.
  librosa.effects.trim(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512)
      rms over centred (reflect-padded) frames -> power_to_db(ref=max, amin=1e-10) > -top_db; keeps
      [first non-silent frame * hop, min(len, (last non-silent frame + 1) * hop)).

PARITY PIN: "unpinned by a live librosa" - there is none to run here.  What pins this file instead: (1) exact algebraic
identities the published algorithm guarantees (istft(stft(y)) == y on the interior to float32 round-off: the Hann-800 / hop-200
pair is COLA for the sum-of-squares normalisation; a pure tone lands in its bin with the window's known gain), checked in
tests/test_dsp_oracle.py, and (2) the reference's own call sites (argument order, constants).  DESIGN.md says so.

Precision note: numpy's FFT runs in float64 and librosa casts the result to complex64 / float32 - restated that way.
"""
import numpy as np
from scipy import signal

# hps/hps.py:17-36
SR = 16000
N_FFT = 1024
HOP = 200            # int(sr * 0.0125)
WIN = 800            # int(sr * 0.05)
N_ITER = 300
PREEMPHASIS = 0.97
MAX_DB = 100
REF_DB = 20


def padded_window(n_fft=N_FFT, win_length=WIN):
    """get_window('hann', win_length, fftbins=True) centred in n_fft samples (librosa.util.pad_center)."""
    w = signal.get_window('hann', win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad)).astype(np.float32)


def stft(y, n_fft=N_FFT, hop=HOP, win_length=WIN):
    """librosa.stft as the reference calls it (preprocess.py:237-240, convert.py:47): (1 + n_fft/2, 1 + len(y)//hop) complex64."""
    y = np.asarray(y, np.float32)
    w = padded_window(n_fft, win_length)
    yp = np.pad(y, n_fft // 2, mode='reflect')
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[idx] * w[None, :]
    return np.fft.rfft(frames, axis=1).T.astype(np.complex64)


def window_sumsquare(n_frames, n_fft=N_FFT, hop=HOP, win_length=WIN):
    """librosa.filters.window_sumsquare(window='hann', norm=None): overlap-added squared (padded) window."""
    w2 = padded_window(n_fft, win_length).astype(np.float32) ** 2
    out = np.zeros(n_fft + hop * (n_frames - 1), np.float32)
    for i in range(n_frames):
        out[i * hop:i * hop + n_fft] += w2
    return out


def istft(S, hop=HOP, win_length=WIN):
    """librosa.istft as the reference calls it (convert.py:42): real signal of hop * (n_frames - 1) samples, float32."""
    S = np.asarray(S)
    n_fft = 2 * (S.shape[0] - 1)
    n_frames = S.shape[1]
    w = padded_window(n_fft, win_length)
    y = np.zeros(n_fft + hop * (n_frames - 1), np.float32)
    frames = (np.fft.irfft(S.T.astype(np.complex128), n=n_fft, axis=1).astype(np.float32)) * w[None, :]
    for i in range(n_frames):
        y[i * hop:i * hop + n_fft] += frames[i]
    wss = window_sumsquare(n_frames, n_fft, hop, win_length)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def griffin_lim(mag, n_iter=N_ITER):
    """convert.py:39-52.  mag: (513, T) linear-amplitude spectrogram."""
    mag = np.asarray(mag, np.float32)
    X_best = mag.astype(np.complex64)                 # copy.deepcopy(spectrogram): zero phase
    for _ in range(n_iter):
        X_t = istft(X_best)
        est = stft(X_t)
        phase = est / np.maximum(1e-8, np.abs(est))
        X_best = (mag * phase).astype(np.complex64)
    return np.real(istft(X_best))


def rms_frames(y, frame_length=2048, hop=512):
    """librosa.feature.rms(y=..., center=True, pad_mode='reflect'): one value per hop."""
    yp = np.pad(np.asarray(y, np.float32), frame_length // 2, mode='reflect')
    n = 1 + (len(yp) - frame_length) // hop
    idx = np.arange(frame_length)[None, :] + hop * np.arange(n)[:, None]
    return np.sqrt(np.mean(np.abs(yp[idx]) ** 2, axis=1))


def trim_bounds(y, top_db=60, frame_length=2048, hop=512):
    """librosa.effects.trim index logic -> (start, end) samples."""
    mse = rms_frames(y, frame_length, hop) ** 2
    amin = 1e-10
    db = 10.0 * np.log10(np.maximum(amin, mse)) - 10.0 * np.log10(np.maximum(amin, mse.max()))
    nz = np.flatnonzero(db > -top_db)
    if nz.size == 0:
        return 0, 0
    return int(nz[0] * hop), int(min(len(y), (nz[-1] + 1) * hop))


def denormalise(mag_norm):
    """convert.py:57-58: clip to [0, 1], back to dB, to linear amplitude."""
    db = np.clip(np.asarray(mag_norm, np.float32), 0, 1) * MAX_DB - MAX_DB + REF_DB
    return np.power(10.0, db * 0.05).astype(np.float32)


def spectrogram2wav(mag_norm, n_iter=N_ITER, trim=True):
    """convert.py:55-62.  mag_norm: (T, 513) normalised log-magnitude spectrogram -> float32 waveform."""
    mag = denormalise(np.asarray(mag_norm).T)
    wav = griffin_lim(mag, n_iter)
    wav = signal.lfilter([1], [1, -PREEMPHASIS], wav)          # de-pre-emphasis
    if trim:
        s, e = trim_bounds(wav)
        wav = wav[s:e]
    return wav.astype(np.float32)


def spectrogram_from_wav(y):
    """preprocess.py:231-256 after `librosa.load` / `effects.trim` (file decoding and resampling stay with the caller):
    y (16 kHz float waveform) -> (T, 513) normalised log-magnitude spectrogram, float32."""
    y = np.asarray(y, np.float32)
    y = np.append(y[0], y[1:] - PREEMPHASIS * y[:-1])           # :233 pre-emphasis
    mag = np.abs(stft(y))                                       # :235-242
    mag = 20 * np.log10(np.maximum(1e-5, mag))                  # :250
    mag = np.clip((mag - REF_DB + MAX_DB) / MAX_DB, 1e-8, 1)    # :254
    return mag.T.astype(np.float32)                             # :258
