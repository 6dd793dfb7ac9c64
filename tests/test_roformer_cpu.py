"""CPU checks of the BS-Roformer side: the oracle against the reference-generated golden and the chunk grid / window of the Roformer branch."""
import os

import numpy as np

import mdx_oracle as M
import roformer_oracle as R

SMALL = dict(dim=32, depth=2, time_transformer_depth=1, freq_transformer_depth=2, freqs_per_bands=(2, 2, 4, 4, 8, 12, 16, 17), dim_head=8, heads=4, stft_n_fft=128,
             stft_hop_length=32, stft_win_length=128, dim_t=65, overlap=8)


def test_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "roformer_small.npz"))
    cfg = R.BSRoformerConfig(**SMALL)
    w = R.make_weights(cfg, seed=3)
    mix = M.synth_music(int(z["n_samples"]), seed=int(z["mix_seed"]))
    y = R.forward(w, cfg, mix[None, :, : cfg.chunk_size])
    assert y.shape == z["forward_ref"].shape and np.abs(y - z["forward_ref"]).max() <= 2e-5
    d = R.demix(mix, cfg, lambda c: R.forward(w, cfg, c), n_instruments=2)
    assert d.shape == (2, 2, mix.shape[1]) and np.abs(d[0] - z["demix_ref"]).max() <= 2e-5
    assert np.array_equal(d[0], d[1])  # a single-target model broadcasts its output into every row of the result tensor (mdxc_separator.py:313)
    cfg_o = R.BSRoformerConfig(**dict(SMALL, overlap=0.03))
    d_o = R.demix(mix, cfg_o, lambda c: R.forward(w, cfg, c), n_instruments=2)
    assert np.abs(d_o[0] - z["demix_overlap_ref"]).max() <= 2e-5


def test_chunk_grid_and_rotary_restatement():
    cfg = R.BSRoformerConfig(**SMALL)
    assert cfg.chunk_size == 32 * 64 and cfg.step == cfg.chunk_size  # overlap 8 s * 44100 > chunk: step clamps to the chunk size (no overlap)
    assert R.BSRoformerConfig(**dict(SMALL, overlap=0.03)).step == 1323
    assert sum(R.DEFAULT_FREQS_PER_BANDS) == 1025 and len(R.DEFAULT_FREQS_PER_BANDS) == 62
    import torch

    fr = torch.from_numpy(R.rotary_freqs(8))
    assert np.allclose(fr.numpy(), 1.0 / 10000 ** (np.arange(0, 8, 2) / 8))
    t = torch.randn(2, 3, 5, 8)
    r = R.apply_rotary(t, fr)
    assert torch.allclose(r[..., 0, :], t[..., 0, :])  # position 0 is not rotated
    assert torch.allclose((r**2).sum(-1), (t**2).sum(-1), atol=1e-5)  # rotations preserve the norm of every pair
    # relative-position property: <rot(q, m), rot(k, n)> depends on m - n only
    q, k = torch.randn(8), torch.randn(8)
    Q, K = R.apply_rotary(q.repeat(6, 1), fr), R.apply_rotary(k.repeat(6, 1), fr)
    assert torch.allclose(Q[1] @ K[0], Q[4] @ K[3], atol=1e-5) and torch.allclose(Q[2] @ K[5], Q[0] @ K[3], atol=1e-5)


def test_melband_oracle_and_band_layout(golden_dir, lib_built):
    z = np.load(os.path.join(golden_dir, "roformer_small.npz"))
    MS = dict(dim=32, depth=2, time_transformer_depth=1, freq_transformer_depth=1, num_bands=12, dim_head=8, heads=4, mask_estimator_depth=2, stft_n_fft=128, stft_hop_length=32,
              stft_win_length=128, dim_t=65)
    cfg = R.MelBandRoformerConfig(**MS)
    w = R.make_mel_weights(cfg, seed=6)
    mix = M.synth_music(int(z["n_samples"]), seed=int(z["mix_seed"]))
    y = R.forward_mel(w, cfg, mix[None, :, : cfg.chunk_size])
    assert np.abs(y - z["mel_forward_ref"]).max() <= 2e-5
    # the product re-derives the band layout on its own (no oracle import on the product path): both must agree bin for bin
    from audio_separator.separator.b200 import roformer as rf

    for kw in (dict(), dict(num_bands=12, stft_n_fft=128, stft_hop_length=32, stft_win_length=128), dict(num_bands=64, sample_rate=48000, stft_n_fft=4096, stft_win_length=4096)):
        a = R.mel_band_layout(R.MelBandRoformerConfig(**kw))
        b = rf.mel_band_layout(rf.MelBandRoformerConfig(**kw))
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
        mask, idx, nfpb, nbpf = b
        assert mask[0, 0] and mask[-1, -1] and nbpf.min() >= 1 and nfpb.sum() * 2 == len(idx)
        assert (np.diff(np.where(mask[5])[0]) == 1).all()  # a band covers a contiguous run of bins


def test_roformer_config_parsing(lib_built):
    """The YAML `model` section -> graph configuration (roformer_loader.py:120-195): defaults, tuple conversion, and loud failures for what is not covered."""
    import pytest

    from audio_separator.separator.b200 import roformer as rf

    c = rf.BSRoformerConfig.from_model_section({"dim": 384, "depth": 6, "stereo": True, "num_stems": 1, "time_transformer_depth": 1, "freq_transformer_depth": 1,
                                                "freqs_per_bands": list(rf.DEFAULT_FREQS_PER_BANDS), "dim_head": 64, "heads": 8, "stft_hop_length": 441, "mask_estimator_depth": 2,
                                                "attn_dropout": 0.1, "flash_attn": True, "multi_stft_hop_size": 147})  # training-only keys are ignored
    assert (c.dim, c.depth, c.stft_hop_length, c.mask_estimator_depth) == (384, 6, 441, 2) and isinstance(c.freqs_per_bands, tuple) and len(c.band_dims) == 62
    assert sum(c.band_dims) == 1025 * 4
    with pytest.raises(ValueError):
        rf.BSRoformerConfig.from_model_section({"dim": 64, "depth": 1, "freqs_per_bands": [2, 2, 4]})  # does not add up to n_fft/2 + 1
    for bad in ({"linear_transformer_depth": 1}, {"sage_attention": True}, {"stft_normalized": True}, {"stft_win_length": 1024}, {"stft_window_fn": "torch.hamming_window"}):
        with pytest.raises(NotImplementedError):
            rf.BSRoformerConfig.from_model_section(dict({"dim": 64, "depth": 1}, **bad))
    m = rf.MelBandRoformerConfig.from_model_section({"dim": 384, "depth": 6, "stereo": True, "num_bands": 60, "stft_hop_length": 441, "mask_estimator_depth": 2, "sample_rate": 44100})
    assert len(m.band_dims) == 60 and sum(m.band_dims) == 2 * 3958 and m.freqs_per_bands[0] == 7
    with pytest.raises(NotImplementedError):
        rf.MelBandRoformerConfig.from_model_section({"dim": 64, "depth": 1, "num_bands": 8, "match_input_audio_length": True})
