"""GPU parity of the Ensembler mirror against golden vectors produced by the UNMODIFIED reference class (oracle/make_golden_ensemble.py)."""
import logging
import os

import numpy as np
import pytest
import torch

import mdx_oracle as M

pytestmark = pytest.mark.gpu


def _waves():
    waves = [M.synth_music(n, seed=60 + i) * g for i, (n, g) in enumerate(((9000, 1.0), (8700, 0.8), (9000, 1.1), (8900, 0.9)))]
    waves[2][:, 100:200] = waves[0][:, 100:200]  # exact ties between models
    return waves


@pytest.mark.parametrize("algo", ["avg_wave", "median_wave", "min_wave", "max_wave", "avg_fft", "median_fft", "min_fft", "max_fft", "uvr_max_spec", "uvr_min_spec"])
def test_ensembler_vs_reference_golden(lib_built, golden_dir, algo):
    assert torch.cuda.is_available()
    from audio_separator.separator.ensembler import Ensembler

    z = np.load(os.path.join(golden_dir, "ensemble_small.npz"))
    waves = _waves()
    log = logging.getLogger("t")
    for tag, wl, wt in (("4", waves, [1.0, 2.0, 0.5, 1.5]), ("3", waves[:3], None)):
        got = Ensembler(log, algo, wt).ensemble([w.copy() for w in wl])
        ref = z[f"{algo}_{tag}"]
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert np.abs(got - ref).max() <= 2e-5, np.abs(got - ref).max()
    if f"{algo}_mono" in z.files:
        got = Ensembler(log, algo).ensemble([w[:1].copy() for w in waves[:2]])
        assert got.shape == z[f"{algo}_mono"].shape and np.abs(got - z[f"{algo}_mono"]).max() <= 2e-5


def test_ensembler_edges(lib_built):
    from audio_separator.separator.ensembler import Ensembler

    log = logging.getLogger("t")
    w = _waves()
    assert Ensembler(log).ensemble([]) is None
    assert Ensembler(log).ensemble([w[0]]) is w[0]
    with pytest.raises(ValueError):
        Ensembler(log, "no_such").ensemble(w[:2])
    with pytest.raises(ValueError):
        Ensembler(log).ensemble([w[0], w[1][:1]])
    with pytest.raises(NotImplementedError):
        Ensembler(log, "ensemble_wav").ensemble(w[:2])
    # weights of the wrong length / summing to zero fall back to equal weights (ensembler.py:33-43)
    a = Ensembler(log, "avg_wave", [1.0, -1.0]).ensemble([w[0].copy(), w[2].copy()])
    b = Ensembler(log, "avg_wave", [3.0]).ensemble([w[0].copy(), w[2].copy()])
    assert np.abs(a - (w[0] + w[2]) / 2).max() <= 1e-6 and np.abs(b - a).max() <= 1e-7
