"""Pin oracle/ensemble_oracle.py against the UNMODIFIED reference Ensembler and write tests/golden/ensemble_small.npz (librosa.stft / istft served by
the oracle's restatements, which oracle/make_golden_vr.py cross-checks against torch)."""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ensemble_oracle as E  # noqa: E402
import mdx_oracle as M  # noqa: E402
import ref_shim  # noqa: E402
import vr_oracle as V  # noqa: E402
from make_golden_vr import check  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.install()
    lib = sys.modules["librosa"]

    def stft(y, n_fft=2048, hop_length=1024):
        y = np.asarray(y)
        return V.stft(y, n_fft, hop_length) if y.ndim == 1 else np.stack([V.stft(c, n_fft, hop_length) for c in y])

    def istft(s, hop_length=1024, length=None, n_fft=None):
        s = np.asarray(s)
        return V.istft(s, hop_length, length) if s.ndim == 2 else np.stack([V.istft(c, hop_length, length) for c in s])

    lib.stft, lib.istft = stft, istft
    ens = ref_shim.ref_module("audio_separator.separator.ensembler")
    waves = [M.synth_music(n, seed=60 + i) * g for i, (n, g) in enumerate(((9000, 1.0), (8700, 0.8), (9000, 1.1), (8900, 0.9)))]
    waves[2][:, 100:200] = waves[0][:, 100:200]  # exact ties between models
    out = {}
    for algo in ("avg_wave", "median_wave", "min_wave", "max_wave", "avg_fft", "median_fft", "min_fft", "max_fft", "uvr_max_spec", "uvr_min_spec"):
        for tag, wl, wt in (("4", waves, [1.0, 2.0, 0.5, 1.5]), ("3", waves[:3], None)):
            ref = ens.Ensembler(logging.getLogger("ref"), algo, wt).ensemble([w.copy() for w in wl])
            got = E.ensemble([w.copy() for w in wl], algo, wt)
            check(f"Ensembler {algo} ({tag} models)", ref, got, 1e-6)
            out[f"{algo}_{tag}"] = np.asarray(ref, dtype=np.float32)
    mono = [w[:1] for w in waves[:2]]
    for algo in ("avg_fft", "uvr_max_spec", "median_wave"):
        ref = ens.Ensembler(logging.getLogger("ref"), algo).ensemble([w.copy() for w in mono])
        check(f"Ensembler {algo} (mono)", ref, E.ensemble([w.copy() for w in mono], algo), 1e-6)
        out[f"{algo}_mono"] = np.asarray(ref, dtype=np.float32)
    # ensemble_wav (spec_utils.ensemble_wav): the reference's own function on the same waveforms
    for tag, wl in (("4", waves), ("3", waves[:3])):
        ref = ens.Ensembler(logging.getLogger("ref"), "ensemble_wav").ensemble([w.copy() for w in wl])
        out[f"ensemble_wav_{tag}"] = np.asarray(ref, dtype=np.float32)
    np.savez_compressed(os.path.join(GOLD, "ensemble_small.npz"), **out)
    print("wrote tests/golden/ensemble_small.npz; oracle pinned: OK")


if __name__ == "__main__":
    main()
