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
    # ensemble_wav is implemented since the multi-model orchestration (golden comparison: tests/test_ensemble_cpu.py); here: it returns one model's rows
    ew = Ensembler(log, "ensemble_wav").ensemble([w[0].copy(), w[2].copy()])
    assert ew.shape == w[0].shape and all(any(np.array_equal(ew[c], x[c]) for x in (w[0], w[2])) for c in range(ew.shape[0]))
    # weights of the wrong length / summing to zero fall back to equal weights (ensembler.py:33-43)
    a = Ensembler(log, "avg_wave", [1.0, -1.0]).ensemble([w[0].copy(), w[2].copy()])
    b = Ensembler(log, "avg_wave", [3.0]).ensemble([w[0].copy(), w[2].copy()])
    assert np.abs(a - (w[0] + w[2]) / 2).max() <= 1e-6 and np.abs(b - a).max() <= 1e-7


def test_separator_multi_model_ensemble_end_to_end(tmp_path):
    """Separator.load_model([a, b]) + separate(): both models run through their plugin, stems are grouped by canonical name and reduced by the Ensembler
    (separator.py:1242-1412).  Expected files = the same reduction applied to what the two single-model runs write."""
    import json
    import wave as wavmod

    import mdx_oracle as O
    from audio_separator.separator import Separator
    from audio_separator.separator.ensembler import Ensembler

    cfg = O.MDXConfig(n_fft=1536, hop_length=256, dim_f=768, dim_t=32, segment_size=32, g=8)
    md = {"compensate": cfg.compensate, "mdx_dim_f_set": cfg.dim_f, "mdx_dim_t_set": 5, "mdx_n_fft_scale_set": cfg.n_fft, "primary_stem": "Vocals"}
    for name, seed in (("UVR-MDX-NET-tinyA", 7), ("UVR-MDX-NET-tinyB", 8)):
        np.savez(tmp_path / f"{name}.npz", **O.make_convtdfnet_weights(cfg, seed=seed, out_gain=0.05))
        (tmp_path / f"{name}.json").write_text(json.dumps(md))
    mix = O.synth_music(30000, seed=3)
    with wavmod.open(str(tmp_path / "song.wav"), "wb") as wf:
        wf.setnchannels(2); wf.setsampwidth(2); wf.setframerate(44100); wf.writeframes((mix.T * 32767).astype("<i2").tobytes())
    params = {"segment_size": cfg.segment_size, "hop_length": cfg.hop_length, "overlap": cfg.overlap, "batch_size": 2}

    def read(path):
        with wavmod.open(str(path)) as wf:
            return (np.frombuffer(wf.readframes(wf.getnframes()), dtype="<i2").astype(np.float32) / 32768.0).reshape(-1, 2).T

    singles = {}
    for name in ("UVR-MDX-NET-tinyA", "UVR-MDX-NET-tinyB"):
        sep = Separator(model_file_dir=str(tmp_path), output_dir=str(tmp_path / name), mdx_params=params)
        sep.load_model(f"{name}.npz")
        singles[name] = {f.split("_(")[1].split(")")[0]: read(tmp_path / name / f) for f in sep.separate(str(tmp_path / "song.wav"))}
    for algo, weights in (("avg_wave", [2.0, 1.0]), ("max_fft", None)):
        out_dir = tmp_path / f"ens_{algo}"
        sep = Separator(model_file_dir=str(tmp_path), output_dir=str(out_dir), mdx_params=params, ensemble_algorithm=algo, ensemble_weights=weights)
        sep.load_model(["UVR-MDX-NET-tinyA.npz", "UVR-MDX-NET-tinyB.npz"])
        files = sep.separate(str(tmp_path / "song.wav"))
        assert sorted(os.path.basename(f) for f in files) == ["song_(Instrumental)_custom_ensemble_tinyA_tinyB.wav", "song_(Vocals)_custom_ensemble_tinyA_tinyB.wav"]
        for f in files:
            stem = os.path.basename(f).split("_(")[1].split(")")[0]
            want = Ensembler(None, algo, weights).ensemble([singles["UVR-MDX-NET-tinyA"][stem], singles["UVR-MDX-NET-tinyB"][stem]])
            want_pcm = O.to_pcm16(np.ascontiguousarray(np.asarray(want).T), 0.9, 0.0).astype(np.int32)
            got_pcm = (read(f).T.reshape(-1) * 32768.0).astype(np.int32)
            assert got_pcm.shape == want_pcm.shape and np.abs(got_pcm - want_pcm).max() <= 1
