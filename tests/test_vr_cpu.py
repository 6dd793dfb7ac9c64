"""CPU checks of the VR side: the oracle against the reference-generated golden, and the host-side filter / resampler design."""
import os

import numpy as np
import pytest

import mdx_oracle as M
import vr_oracle as V


def test_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "vr_small.npz"))
    for arch, bins, width in ((31191, 128, 272), (129605, 128, 288)):
        w = V.make_weights(arch, seed=arch % 97)
        cfg = V.VRConfig(param=V.single_band_param(n_fft=bins * 2, hl=bins // 2, bins=bins), nn_architecture=arch, window_size=width)
        m = V.predict_mask(w, cfg, z[f"mask_in_{arch}"])
        assert m.shape == z[f"mask_ref_{arch}"].shape and np.abs(m - z[f"mask_ref_{arch}"]).max() <= 2e-6
    arch = 31191
    w = V.make_weights(arch, seed=arch % 97)
    cfg = V.VRConfig(param=V.single_band_param(n_fft=256, hl=64, bins=128), nn_architecture=arch, window_size=272)
    wave = M.synth_music(int(z["n_samples"]), seed=int(z["wave_seed"]))
    assert np.abs(V.loading_mix(wave, cfg) - z["X_1band"]).max() <= 1e-4
    prim, sec = V.separate_arrays(wave, cfg, lambda b: V.predict_mask(w, cfg, b), batch_size=3)
    assert prim.shape == z["prim_1band"].shape == (2, 64 * (z["X_1band"].shape[2] - 1))  # NOT the input length
    assert np.abs(prim - z["prim_1band"]).max() <= 2e-5 and np.abs(sec - z["sec_1band"]).max() <= 2e-5
    cfg4 = V.VRConfig(param=V.four_band_v2_param(), nn_architecture=33966, window_size=272)
    wave4 = M.synth_music(int(z["n_samples4"]), seed=int(z["wave4_seed"]))
    assert np.abs(V.loading_mix(wave4, cfg4) - z["X_4band"]).max() <= 1e-4


def test_make_padding_and_aggressiveness_edges():
    assert V.make_padding(1000, 512, 128) == (128, 256 - 1000 % 256 + 128, 256)
    assert V.make_padding(512, 512, 128) == (128, 256 + 128, 256)  # an exact multiple still gets one whole extra roi
    assert V.make_padding(10, 256, 128)[2] == 256  # roi_size 0 falls back to the crop size
    m = np.full((2, 10, 3), 0.5, np.float32)
    out = V.adjust_aggr(m.copy(), False, 0.1, 4, {"left": 0.3, "right": 0.0})
    assert np.allclose(out[0, :4], 0.5 ** (1 + 0.5 / 3)) and np.allclose(out[0, 4:], 0.5**1.5) and np.allclose(out[1, 4:], 0.5**1.2)
    assert np.array_equal(V.adjust_aggr(m.copy(), True, 0.0, 4), m)  # aggression 0: untouched even for a vocal primary stem
    assert np.allclose(V.adjust_aggr(m.copy(), True, 0.1, 4)[0, 4:], 0.5**1.8)  # non-accompaniment stems use 1 - aggr


def test_resampler_design_and_filters_match_scipy_and_the_oracle(lib_built):
    import scipy.signal

    from audio_separator.separator.b200 import vr, vr_params

    for up, down in ((1, 3), (1, 2), (2, 1), (3, 1), (2, 3)):
        taps, pre, u, d = vr.resample_poly_design(up, down)
        half = 10 * max(up, down)
        h = scipy.signal.firwin(2 * half + 1, 1.0 / max(up, down), window=("kaiser", 5.0)) * up
        n_pre_pad = down - half % down
        assert len(taps) == n_pre_pad + 2 * half + 1 and not taps[:n_pre_pad].any()
        assert np.abs(taps[n_pre_pad:] - h).max() <= 1e-7
        x = np.random.default_rng(up * 7 + down).standard_normal(1001)
        ref = scipy.signal.resample_poly(x, up, down)
        n_out = -(-len(x) * up // down)
        got = np.array([sum(x[i] * float(taps[(k + pre) * down - i * up]) for i in range(max(0, -(-((k + pre) * down - len(taps) + 1) // up)), min(len(x) - 1, (k + pre) * down // up) + 1))
                        for k in range(n_out)])
        assert len(ref) == n_out and np.abs(got - ref).max() <= 1e-6
    spec = np.ones((2, 40, 3), complex)
    assert np.allclose(V.fft_lp_filter(spec.copy(), 25, 33)[0, :, 0], vr.lp_gain(40, 25, 33))
    assert np.allclose(V.fft_hp_filter(spec.copy(), 20, 9)[0, :, 0], vr.hp_gain(40, 20, 9))
    p = vr_params.load("4band_v2", "/nonexistent")
    ref = V.four_band_v2_param()
    assert p["bins"] == ref["bins"] and p["band"][3] == ref["band"][3] and p["reverse"] is False
    with pytest.raises(FileNotFoundError):
        vr_params.load("no_such_layout", "/nonexistent")


def test_vr51_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "vr_small.npz"))
    w = V.make_weights_51(256, 8, 16, seed=3)
    m = V.predict_mask_51(w, 256, z["mask51_in"])
    assert np.abs(m - z["mask51_ref"]).max() <= 5e-6
    p51 = V.single_band_param(n_fft=256, hl=64, bins=128, pre_filter_start=120, pre_filter_stop=127)
    p51["band"][1]["convert_channels"] = "mid_side_c"
    cfg = V.VRConfig(param=p51, window_size=160, aggression=5, primary_stem="Vocals", offset=64, is_51=True, nout=8, nout_lstm=16)
    wave = M.synth_music(int(z["n_samples"]), seed=int(z["wave_seed"]))
    assert np.abs(V.loading_mix(wave, cfg) - z["X_51"]).max() <= 1e-4
    # filter masks of the 5.1 glue: length and end points (spec_utils.py:398-407)
    lp, hp = V.lp_filter_mask(100, 40, 60)[:, 0], V.hp_filter_mask(100, 30, 10)[:, 0]
    assert len(lp) == 100 and lp[38] == 1 and lp[39] == 1 and lp[59] == 0 and lp[60] == 0 and 0 < lp[50] < 1
    assert len(hp) == 100 and hp[10] == 0 and hp[11] == 0 and hp[31] == 1 and 0 < hp[20] < 1


def test_post_process_run_logic_matches_the_oracle(lib_built, golden_dir):
    """merge_artifacts: the oracle is pinned against the reference function (oracle/make_golden_vr.py); the product's host-side run detection
    must give the same masks on random run layouts (runs at the start / end, runs closer than one fade, runs that are too short, no run at all)."""
    from audio_separator.separator.b200 import vr

    z = np.load(os.path.join(golden_dir, "vr_small.npz"))
    assert np.array_equal(V.merge_artifacts(z["pp_mask_in"].copy(), thres=0.2), z["pp_mask_ref"])
    rng = np.random.default_rng(0)
    for _ in range(120):
        n = int(rng.integers(100, 1200))
        mk = rng.uniform(0, 0.15, (2, 5, n)).astype(np.float32)
        pos = 0
        while pos < n - 5:
            pos += int(rng.integers(1, 150))
            if rng.random() < 0.15:
                pos = 0
            L = int(rng.integers(5, 300))
            mk[:, :, pos : pos + L] = 0.7
            pos += L
        if rng.random() < 0.2:
            mk[:, :, -int(rng.integers(70, 200)) :] = 0.8
        ref = V.merge_artifacts(mk.copy(), thres=0.2)
        w = vr.merge_weights(mk.min(axis=(0, 1)), 0.2)
        got = mk.copy() if w is None else mk + w[None, None, :] * (1 - mk)
        assert np.array_equal(got, ref)


def test_band_upsampler_against_libsamplerate_when_present():
    """the multi-band synthesis up-samples with librosa.resample(res_type="sinc_fastest") = libsamplerate, which is in neither the
    reference tree nor this image, so oracle and product share a Kaiser polyphase stand-in and that ONE step is "parity unpinned".  Wherever the
    `samplerate` wheel exists (a box with the reference installed) this test pins it: the stand-in must stay within the audio gate of libsamplerate
    on band-limited material, and it fails loudly if it does not -- telling the maintainer the stand-in has to be replaced by the exact table."""
    samplerate = pytest.importorskip("samplerate")
    y = M.synth_music(7350 * 4, seed=3)[:, ::6].astype(np.float32)  # 4 s at 7350 Hz, the lowest 4band_v2 band rate
    for sr_in, sr_out in ((7350, 14700), (14700, 44100)):
        ref = np.stack([samplerate.resample(ch, sr_out / sr_in, "sinc_fastest") for ch in y])
        got = V.upsample(y, sr_in, sr_out)
        n = min(ref.shape[1], got.shape[1])
        core = slice(64, n - 64)  # the two resamplers treat the edges differently
        assert np.abs(got[:, core] - ref[:, core]).max() <= 1e-4
