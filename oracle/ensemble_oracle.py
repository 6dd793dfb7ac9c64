"""CPU oracle for the Ensembler (audio_separator/separator/ensembler.py:10-156; spec_utils.ensembling :583-608, wave_to_spectrogram_no_mp /
spectrogram_to_wave_no_mp :538-554) -- TEST INFRASTRUCTURE.  librosa.stft / istft are the restatements of vr_oracle.py.
Pinned against the unmodified reference class by oracle/make_golden_ensemble.py.  `ensemble_wav` is not restated."""
import numpy as np

import vr_oracle as V


def _stft2(w):
    return np.stack([V.stft(w[0], 2048, 1024), V.stft(w[1], 2048, 1024)])


def _pick(arr, take_max):
    """Ensembler._lambda_max / _lambda_min with key=np.abs, axis=0: the first extremum wins."""
    k = np.abs(arr)
    idx = np.argmax(k, 0) if take_max else np.argmin(k, 0)
    return np.take_along_axis(arr, idx[None], 0)[0]


def ensemble(waveforms, algorithm="avg_wave", weights=None):
    if not waveforms:
        return None
    if len(waveforms) == 1:
        return waveforms[0]
    nch = waveforms[0].shape[0]
    L = max(w.shape[1] for w in waveforms)
    ws = [np.pad(w, ((0, 0), (0, L - w.shape[1]))) if w.shape[1] < L else w for w in waveforms]
    wt = np.ones(len(ws)) if weights is None else np.array(weights)
    if len(wt) != len(ws) or not np.all(np.isfinite(wt)) or not np.isfinite(wt.sum()) or wt.sum() == 0:
        wt = np.ones(len(ws))
    if algorithm == "avg_wave":
        out = np.zeros_like(ws[0])
        for w, a in zip(ws, wt):
            out += w * a
        return out / np.sum(wt)
    if algorithm == "median_wave":
        return np.median(ws, axis=0)
    if algorithm in ("min_wave", "max_wave"):
        return _pick(np.array(ws), algorithm == "max_wave")
    if algorithm in ("avg_fft", "median_fft", "min_fft", "max_fft"):
        specs = np.array([_stft2(w if nch == 2 else np.vstack([w, w])) for w in ws])
        if algorithm == "avg_fft":
            e = np.zeros_like(specs[0])
            for s_, a in zip(specs, wt):
                e += s_ * a
            e /= np.sum(wt)
        elif algorithm == "median_fft":
            e = np.median(np.real(specs), axis=0) + 1j * np.median(np.imag(specs), axis=0)
        else:
            e = _pick(specs, algorithm == "max_fft")
        wave = np.stack([V.istft(e[0], 1024, length=L), V.istft(e[1], 1024, length=L)])
        return wave[:1] if nch == 1 else wave
    if algorithm in ("uvr_max_spec", "uvr_min_spec"):
        specs = [np.stack([V.stft(c, 2048, 1024) for c in w]) for w in ws]
        cur = specs[0]
        for s_ in specs[1:]:
            ln = min(cur.shape[2], s_.shape[2])
            cur, s_ = cur[:, :, :ln], s_[:, :, :ln]
            cur = np.where(np.abs(s_) >= np.abs(cur), s_, cur) if algorithm == "uvr_max_spec" else np.where(np.abs(s_) <= np.abs(cur), s_, cur)
        return np.stack([V.istft(c, 1024) for c in cur])
    raise ValueError(f"Unknown ensemble algorithm: {algorithm}")
