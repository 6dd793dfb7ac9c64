"""The ensembler oracle against the reference-generated golden (CPU)."""
import os

import numpy as np

import ensemble_oracle as E
import mdx_oracle as M


def test_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "ensemble_small.npz"))
    waves = [M.synth_music(n, seed=60 + i) * g for i, (n, g) in enumerate(((9000, 1.0), (8700, 0.8), (9000, 1.1), (8900, 0.9)))]
    waves[2][:, 100:200] = waves[0][:, 100:200]
    for algo in ("avg_wave", "median_wave", "min_wave", "max_wave", "avg_fft", "median_fft", "max_fft", "uvr_min_spec"):
        got = E.ensemble([w.copy() for w in waves], algo, [1.0, 2.0, 0.5, 1.5])
        assert got.shape == z[f"{algo}_4"].shape and np.abs(got - z[f"{algo}_4"]).max() <= 1e-6
    assert E.ensemble([], "avg_wave") is None and E.ensemble([waves[0]], "max_fft") is waves[0]
    # spectral results keep the (padded) input length; the UVR variants return hop * (frames - 1) samples
    assert z["avg_fft_4"].shape == (2, 9000) and z["uvr_max_spec_4"].shape == (2, 1024 * (9000 // 1024))


def test_ensemble_wav_mirror_matches_reference_golden(golden_dir, lib_built):
    """Ensembler(algorithm="ensemble_wav") is host arithmetic in the mirror too (a per-channel pick of the quietest model): compare with the reference's output."""
    from audio_separator.separator.ensembler import Ensembler

    z = np.load(os.path.join(golden_dir, "ensemble_small.npz"))
    waves = [M.synth_music(n, seed=60 + i) * g for i, (n, g) in enumerate(((9000, 1.0), (8700, 0.8), (9000, 1.1), (8900, 0.9)))]
    waves[2][:, 100:200] = waves[0][:, 100:200]
    for tag, wl in (("4", waves), ("3", waves[:3])):
        got = Ensembler(None, "ensemble_wav").ensemble([w.copy() for w in wl])
        assert got.shape == z[f"ensemble_wav_{tag}"].shape and np.array_equal(got.astype(np.float32), z[f"ensemble_wav_{tag}"])
