"""CPU checks of the Demucs side: the oracle against the reference-generated golden, and the host-side package / bag loader."""
import os
import sys
import types
from fractions import Fraction

import numpy as np
import pytest
import torch

import demucs_oracle as D
import mdx_oracle as M

SMALL = dict(channels=8, bottom_channels=32, t_layers=3, t_heads=4, segment=Fraction(1, 2))


def test_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "demucs_small.npz"))
    cfg = D.HTConfig(**SMALL)
    w = D.make_weights(cfg, seed=int(z["weights_seed"]))
    mix = M.synth_music(3 * cfg.seg_len, seed=int(z["mix_seed"]))
    y = D.forward(w, cfg, mix[None, :, : cfg.seg_len])
    # torch-CPU reductions are partitioned by thread count: seen up to ~2e-5 between an idle and a loaded host, so the gate is 5e-5 (1e-4 on audio is the contract)
    assert np.abs(y - z["forward_ref"]).max() <= 5e-5
    ys = D.forward(w, cfg, mix[None, :, : cfg.seg_len - 1234])
    assert ys.shape == z["forward_short_ref"].shape and np.abs(ys - z["forward_short_ref"]).max() <= 5e-5
    N = int(z["n_apply"])
    src = D.demix_demucs([lambda c: D.forward(w, cfg, c)], [[1.0] * 4], cfg, mix[:, :N], [[int(v) for v in z["shift_offsets"]]], 0.25)
    assert src.shape == (4, 2, N) and np.abs(src - z["demix_ref"]).max() <= 5e-5


def test_center_trim_and_padded_edges():
    a = np.arange(10.0)[None]
    assert D.center_trim(a, 10) is a
    assert D.center_trim(a, 7).tolist() == [[1, 2, 3, 4, 5, 6, 7]]  # delta 3: 1 off the left, 2 off the right (utils.py:53-70)
    t = np.arange(1.0, 6.0)[None]
    assert D.padded(t, 3, 2, 6).tolist() == [[2, 3, 4, 5, 0, 0]]  # real neighbours on the left, zeros past the end (apply.py:97-113)
    assert D.padded(t, 0, 2, 6).tolist() == [[0, 0, 1, 2, 3, 4]]


def _fake_package(path, kwargs, state, klass_name="HTDemucs"):
    mod, parent = types.ModuleType("demucs_fake_pkg.htdemucs"), types.ModuleType("demucs_fake_pkg")
    K = type(klass_name, (), {"__module__": "demucs_fake_pkg.htdemucs"})
    setattr(mod, klass_name, K)
    sys.modules["demucs_fake_pkg"], sys.modules["demucs_fake_pkg.htdemucs"] = parent, mod
    try:
        torch.save({"klass": K, "args": (), "kwargs": kwargs, "state": state}, path)
    finally:
        del sys.modules["demucs_fake_pkg"], sys.modules["demucs_fake_pkg.htdemucs"]


def test_package_and_bag_loader(lib_built, tmp_path):
    from audio_separator.separator.b200 import demucs_loader as L

    cfg = D.HTConfig(**SMALL)
    w = D.make_weights(cfg, seed=1)
    state = {k: torch.from_numpy(v).half() for k, v in w.items()}  # released packages store half precision
    _fake_package(str(tmp_path / "abcd0001-deadbeef.th"), cfg.kwargs(), state)
    _fake_package(str(tmp_path / "abcd0002.th"), cfg.kwargs(), state)
    (tmp_path / "bag.yaml").write_text("models: ['abcd0001', 'abcd0002']\nweights: [[1, 0, 0, 0], [0, 1, 1, 1]]\nsegment: 4\n")
    models, weights, segment = L.load_demucs(str(tmp_path / "bag.yaml"))
    assert len(models) == 2 and weights == [[1, 0, 0, 0], [0, 1, 1, 1]] and segment == 4
    c, st = models[0]
    assert (c.channels, c.bottom_channels, c.t_layers, c.t_heads, c.segment, c.sources) == (8, 32, 3, 4, Fraction(1, 2), ("drums", "bass", "other", "vocals"))
    assert set(st) == set(w) and all(v.dtype == np.float32 for v in st.values())
    assert np.abs(st["encoder.0.conv.weight"] - w["encoder.0.conv.weight"]).max() <= 2e-3  # half-precision storage
    single, wts, seg = L.load_demucs(str(tmp_path / "abcd0002.th"))
    assert len(single) == 1 and wts is None and seg is None
    # constructor defaults apply when the package leaves an argument out (htdemucs.py:56-133)
    _fake_package(str(tmp_path / "defaults.th"), {"sources": ["a", "b"]}, state)
    c2 = L.config_from_package(L.load_package(str(tmp_path / "defaults.th")))
    assert (c2.bottom_channels, c2.segment, c2.channels, c2.sources) == (0, Fraction(10), 48, ("a", "b"))
    # anything outside the supported structure fails loudly instead of producing different audio
    _fake_package(str(tmp_path / "v2.th"), {}, state, klass_name="Demucs")
    with pytest.raises(NotImplementedError):
        L.config_from_package(L.load_package(str(tmp_path / "v2.th")))
    _fake_package(str(tmp_path / "sparse.th"), dict(cfg.kwargs(), t_sparse_self_attn=True), state)
    with pytest.raises(NotImplementedError):
        L.config_from_package(L.load_package(str(tmp_path / "sparse.th")))
    (tmp_path / "broken.yaml").write_text("models: ['nope']\n")
    with pytest.raises(FileNotFoundError):
        L.load_demucs(str(tmp_path / "broken.yaml"))


def test_host_side_weight_blocking_and_embeddings(lib_built):
    from audio_separator.separator.b200 import demucs as dm

    rng = np.random.default_rng(0)
    w = rng.standard_normal((50, 7, 3, 3)).astype(np.float32)
    b = dm.block_conv_weight(w)
    assert b.shape == (7, 9, 96) and np.array_equal(b[3, 4, :50], w[:, 3, 1, 1]) and not b[:, :, 50:].any()
    wt = rng.standard_normal((6, 5, 8)).astype(np.float32)
    bt = dm.block_convtr_weight(wt, 4)
    assert bt.shape == (6, 2, 48)
    assert np.array_equal(bt[2, 0, 3 * 5 : 4 * 5], wt[2, :, 7]) and np.array_equal(bt[2, 1, 0:5], wt[2, :, 0])
    pe = torch.from_numpy(D.sin_embedding_2d(32, 8, 22)).permute(2, 1, 0).reshape(22 * 8, 32).numpy()
    assert np.abs(pe - dm.sin_embedding_2d_tokens(32, 8, 22, 10000.0)).max() <= 1e-6
    assert np.abs(D.sin_embedding_1d(87, 32) - dm.sin_embedding_1d(87, 32, 10000.0)).max() <= 1e-6
    with pytest.raises(ValueError):
        dm.HTDemucsConfig(kernel_size=4).validate()


# ----------------------------------------------------------------------------------------------------------------- Hybrid Demucs v3
HD_SMALL = dict(channels=8, nfft=256, depth=4, norm_starts=2, dconv_lstm=2, dconv_attn=2, segment=0.5)


def test_hdemucs_oracle_matches_reference_golden(golden_dir):
    import hdemucs_oracle as H

    z = np.load(os.path.join(golden_dir, "hdemucs_small.npz"))
    cfg = H.HDConfig(**HD_SMALL)
    w = H.make_weights(cfg, seed=int(z["weights_seed"]))
    mix = M.synth_music(3 * cfg.seg_len, seed=int(z["mix_seed"]))
    L = int(z["seg_len"])
    assert np.abs(H.forward(w, cfg, mix[None, :, :L]) - z["forward_ref"]).max() <= 5e-5
    yl = H.forward(w, cfg, mix[None, :, : int(z["long_len"])])  # ragged length, BLSTM input framed into overlapping 200-step windows
    assert yl.shape == z["forward_long_ref"].shape and np.abs(yl - z["forward_long_ref"]).max() <= 5e-5
    yo = H.forward(w, H.HDConfig(**dict(HD_SMALL, hybrid_old=True)), mix[None, :, :L])
    assert np.abs(yo - z["forward_old_ref"]).max() <= 5e-5
    N = int(z["n_apply"])
    m2 = torch.from_numpy(mix[:, :N])
    mn = ((m2 - m2.mean(0).mean()) / m2.mean(0).std()).numpy()
    a = H.apply_model(lambda c: H.forward(w, cfg, c), cfg, mn[None], [int(v) for v in z["shift_offsets"]], 0.25)
    assert np.abs(a - z["apply_ref"]).max() <= 1e-4 * np.abs(z["apply_ref"]).max()


def test_hdemucs_layer_plan_and_package_loader(lib_built, tmp_path):
    """The layer geometry of the released hybrid models (depth 6, nfft 4096) and HDemucs packages through the loader."""
    import hdemucs_oracle as H
    from audio_separator.separator.b200 import demucs_loader as L
    from audio_separator.separator.b200 import hdemucs as hd

    plan = hd.layer_plan(hd.HDemucsConfig())
    assert plan == H.layer_plan(H.HDConfig())  # the product's restatement of the constructor agrees with the oracle's (pinned: its state_dict layout loads into the reference)
    e = [p["enc"] for p in plan]
    assert [(l["chin"], l["chout"], l["k"], l["s"], l["freq"], l["pad"], l["norm"], l["lstm"]) for l in e] == [
        (4, 48, 8, 4, True, 2, False, False), (48, 96, 8, 4, True, 2, False, False), (96, 192, 8, 4, True, 2, False, False), (192, 384, 8, 4, True, 2, False, False),
        (384, 768, 8, 4, True, 0, True, True), (768, 1536, 4, 2, False, 1, True, True)]
    assert [p["tenc"]["empty"] for p in plan[:5]] == [False, False, False, False, True] and plan[5]["tenc"] is None
    assert plan[0]["dec"]["chout"] == 16 and plan[0]["tdec"]["chout"] == 8 and plan[0]["dec"]["last"]
    cfg = H.HDConfig(**HD_SMALL)
    w = H.make_weights(cfg, seed=2)
    state = {k: torch.from_numpy(v) for k, v in w.items()}
    _fake_package(str(tmp_path / "hd.th"), dict(cfg.kwargs(), cac=True, hybrid=True, wiener_iters=0, multi_freqs=[], rescale=0.1, dconv_init=1e-3), state, klass_name="HDemucs")
    models, wts, seg = L.load_demucs(str(tmp_path / "hd.th"))
    c, st = models[0]
    assert isinstance(c, hd.HDemucsConfig) and (c.channels, c.nfft, c.depth, c.norm_starts, c.dconv_lstm, c.segment, c.pads_to_segment) == (8, 256, 4, 2, 2, 0.5, False)
    assert c.seg_len == 22050 and set(st) == set(w)
    for bad in (dict(multi_freqs=[0.5]), dict(cac=False), dict(hybrid=False), dict(wiener_iters=2)):
        _fake_package(str(tmp_path / "bad.th"), dict(cfg.kwargs(), **bad), state, klass_name="HDemucs")
        with pytest.raises(NotImplementedError):
            L.config_from_package(L.load_package(str(tmp_path / "bad.th")))
    with pytest.raises(ValueError):
        hd.HDemucsConfig(kernel_size=6).validate()


def test_group_units_for_models_with_and_without_valid_length(lib_built):
    """Which segments of an apply_model pass share a forward: everything at the training segment for HTDemucs, runs of equal length for HDemucs."""
    from audio_separator.separator.b200.demucs import group_units

    assert group_units([100, 100, 100, 37], 100, True) == [(0, 4, 100)]
    assert group_units([100, 100, 100, 37], 100, False) == [(0, 3, 100), (3, 4, 37)]
    assert group_units([37], 100, False) == [(0, 1, 37)] and group_units([37], 100, True) == [(0, 1, 100)]
    assert group_units([], 100, True) == [] and group_units([], 100, False) == []
    assert group_units([100, 60, 60, 100], 100, False) == [(0, 1, 100), (1, 3, 60), (3, 4, 100)]
