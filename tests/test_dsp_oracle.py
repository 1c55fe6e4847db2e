"""CPU tests of oracle/dsp_oracle.py (numpy restatement of the librosa calls the reference makes in preprocess.py:231-256 and
convert.py:39-62).  librosa is not installed here and not vendored by the reference, so the restatement is held to the
identities its published algorithm guarantees (see the module docstring: "parity unpinned by a live librosa")."""
import numpy as np
import pytest
from scipy import signal

from oracle import dsp_oracle as dsp


def speechlike(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / dsp.SR
    f0 = 120 + 40 * np.sin(2 * np.pi * 1.3 * t)
    y = sum(np.sin(2 * np.pi * np.cumsum(f0 * h) / dsp.SR) / h for h in range(1, 12))
    env = 0.5 + 0.5 * np.sin(2 * np.pi * 3.1 * t) ** 2
    return (0.05 * y * env + 0.002 * rng.standard_normal(n)).astype(np.float32)


def test_window_is_periodic_hann_centred():
    w = dsp.padded_window()
    assert w.shape == (1024,) and np.all(w[:112] == 0) and np.all(w[912:] == 0)
    assert w[112] == 0 and abs(w[112 + 400] - 1.0) < 1e-7                    # periodic: w[0] = 0, peak at win_length / 2
    assert np.allclose(w[112:912], signal.get_window('hann', 800, fftbins=True))


def test_stft_shape_and_tone_gain():
    n = 16000
    y = np.cos(2 * np.pi * (100 * dsp.SR / 1024) * np.arange(n) / dsp.SR).astype(np.float32)     # exactly bin 100
    S = dsp.stft(y)
    assert S.shape == (513, 1 + n // 200) and S.dtype == np.complex64
    mid = np.abs(S[:, 10:-10])
    assert (mid.argmax(0) == 100).all()
    assert np.allclose(mid[100], dsp.padded_window().sum() / 2, rtol=1e-4)       # a unit cosine: half the window's sum


def test_istft_inverts_stft():
    y = speechlike(12345)
    S = dsp.stft(y)
    yr = dsp.istft(S)
    assert yr.shape == (200 * (S.shape[1] - 1),) and yr.dtype == np.float32
    assert np.abs(yr - y[:len(yr)]).max() < 2e-6                                  # COLA under the sum-of-squares normalisation


def test_stft_matches_scipy_on_the_interior():
    """scipy.signal.stft with the same window / hop / zero-padded frames agrees away from the edges (different edge
    conventions), which cross-checks framing, window placement and scaling."""
    y = speechlike(8000, 1)
    S = dsp.stft(y)
    w = signal.get_window('hann', 800, fftbins=True)
    _, _, Z = signal.stft(y, window=w, nperseg=800, noverlap=600, nfft=1024, boundary=None, padded=False, scaling='spectrum')
    Z = Z * w.sum()                                                                # undo scipy's spectrum scaling
    # scipy frame j starts at sample 200 j (window sample 0); ours at 200 j - 400: ours[j + 2] == scipy[j] up to the phase of the
    # 112-sample offset of the window inside the 1024-sample frame
    k = np.arange(513)[:, None]
    rot = np.exp(-2j * np.pi * k * 112 / 1024)
    n = min(Z.shape[1], S.shape[1] - 2)
    assert np.abs(S[:, 2:2 + n] - Z[:, :n] * rot).max() < 2e-4 * np.abs(S).max()


def test_griffin_lim_reduces_inconsistency():
    y = speechlike(6000, 2)
    mag = np.abs(dsp.stft(y))

    def err(n_iter):
        w = dsp.griffin_lim(mag, n_iter)
        return np.linalg.norm(np.abs(dsp.stft(w)) - mag) / np.linalg.norm(mag)
    e0, e10 = err(0), err(10)
    assert e10 < 0.6 * e0


def test_trim_bounds():
    y = np.concatenate([np.zeros(5000, np.float32), speechlike(8000, 3), np.zeros(7000, np.float32)])
    s, e = dsp.trim_bounds(y)
    assert 3000 <= s <= 5000 and 13000 <= e <= 15000 and s % 512 == 0
    assert dsp.trim_bounds(np.zeros(4000, np.float32)) == (0, 4000)     # all frames equal the max: nothing is 60 dB below


def test_feature_roundtrip_ranges():
    y = speechlike(9000, 4)
    m = dsp.spectrogram_from_wav(y)
    assert m.shape == (1 + 9000 // 200, 513) and m.dtype == np.float32
    assert m.min() >= 1e-8 and m.max() <= 1.0
    w = dsp.spectrogram2wav(m, n_iter=3, trim=False)
    assert w.shape == (200 * (m.shape[0] - 1),) and np.isfinite(w).all()
