"""Pin oracle/vr_oracle.py against the UNMODIFIED reference (vr_network/nets.py + layers.py, spec_utils.py, VRSeparator methods)
and write tests/golden/vr_small.npz.  librosa is absent here: the reference's calls into it are served by the oracle's
restatements (stft / istft cross-checked against torch.stft / torch.istft below; polyphase = scipy.signal.resample_poly);
the libsamplerate up-sampling of the multi-band inverse has no reference available (oracle: stand-in, parity unpinned)."""
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mdx_oracle as M  # noqa: E402
import ref_shim  # noqa: E402
import vr_oracle as V  # noqa: E402


def check(name, ref, got, tol):
    ref, got = np.asarray(ref), np.asarray(got)
    assert ref.shape == got.shape, (name, ref.shape, got.shape)
    err = float(np.abs(ref.astype(np.complex128) - got.astype(np.complex128)).max())
    print(f"  pin {name:<58s} max|ref-oracle| = {err:.3e} (ref max {np.abs(ref).max():.3e}) tol {tol:.1e}")
    if not err <= tol:
        raise SystemExit(f"oracle pin FAILED: {name}")


GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def librosa_stub():
    m = sys.modules["librosa"]
    m.stft = lambda y, n_fft, hop_length: V.stft(y, n_fft, hop_length)
    m.istft = lambda s, hop_length: V.istft(s, hop_length)

    def resample(y, orig_sr, target_sr, res_type="soxr_hq"):
        if res_type == "polyphase":
            return V.resample_polyphase(y, orig_sr, target_sr)
        return V.upsample(y, orig_sr, target_sr)  # sinc_fastest: no reference implementation in this container

    m.resample = resample
    return m


def ref_net(arch, n_fft_bins, w):
    nets = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.vr_network.nets")
    net = nets.determine_model_capacity(n_fft_bins, arch).eval()
    sd = net.state_dict()
    names = [n for n, _ in V.param_shapes(arch)]
    assert list(sd) == names, [(a, b) for a, b in zip(sd, names) if a != b][:5]
    for (n, s), v in zip(V.param_shapes(arch), sd.values()):
        assert tuple(v.shape) == tuple(s), (n, v.shape, s)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w.items()})
    return net


def ref_separator(cfg, net, wave):
    vs = ref_shim.ref_module("audio_separator.separator.architectures.vr_separator")
    mp = types.SimpleNamespace(param=cfg.param)
    sep = object.__new__(vs.VRSeparator)
    sep.logger = logging.getLogger("ref")
    sep.model_params, sep.is_vr_51_model, sep.model_run = mp, False, net
    sep.enable_tta = sep.enable_post_process = sep.high_end_process = False
    sep.batch_size, sep.window_size = 2, cfg.window_size
    sep.input_high_end_h = sep.input_high_end = None
    sep.primary_stem_name = cfg.primary_stem
    sep.torch_device_mps, sep.audio_file_path, sep.wav_subtype = None, "in-memory.wav", "PCM_16"
    sep.aggression = float(int(cfg.aggression) / 100)
    sep.aggressiveness = {"value": sep.aggression, "split_bin": cfg.param["band"][1]["crop_stop"], "aggr_correction": cfg.param.get("aggr_correction")}
    vs.librosa.load = lambda f, sr, mono, dtype, res_type: (wave.astype(np.float32), sr)
    return sep


def main():
    ref_shim.install()
    librosa_stub()
    # ---- the restated librosa transforms against torch's (independent implementation of the same definition)
    rng = np.random.default_rng(0)
    y = rng.standard_normal(5000).astype(np.float32)
    for n_fft, hop in ((960, 480), (512, 160), (320, 80), (640, 80), (2048, 512)):
        win = torch.hann_window(n_fft)
        t = torch.stft(torch.from_numpy(y), n_fft, hop, window=win, center=True, pad_mode="constant", return_complex=True).numpy()
        s = V.stft(y, n_fft, hop)
        check(f"stft n_fft={n_fft} hop={hop} vs torch", t, s, 2e-4)
        ti = torch.istft(torch.from_numpy(s), n_fft, hop, window=win, center=True).numpy()
        check(f"istft n_fft={n_fft} hop={hop} vs torch", ti, V.istft(s, hop), 2e-5)
    out = {}
    # ---- network: three structural variants at reduced size
    for arch, bins, width in ((31191, 128, 272), (33966, 64, 272), (129605, 128, 288), (123821, 64, 272)):
        w = V.make_weights(arch, seed=arch % 97)
        net = ref_net(arch, bins * 2, w)
        x = np.abs(rng.standard_normal((2, 2, bins + 1, width))).astype(np.float32)
        with torch.no_grad():
            m_ref = net.predict_mask(torch.from_numpy(x)).numpy()
        cfg = V.VRConfig(param=V.single_band_param(n_fft=bins * 2, hl=bins // 2, bins=bins), nn_architecture=arch, window_size=width)
        m_orc = V.predict_mask(w, cfg, x)
        check(f"predict_mask arch {arch}", m_ref, m_orc, 2e-6)
        if arch in (31191, 129605):
            out[f"mask_in_{arch}"], out[f"mask_ref_{arch}"] = x, m_ref
    # ---- the whole VRSeparator path, single band (no resampling anywhere: fully pinned)
    arch = 31191
    w = V.make_weights(arch, seed=arch % 97)
    p1 = V.single_band_param(n_fft=256, hl=64, bins=128)
    cfg1 = V.VRConfig(param=p1, nn_architecture=arch, window_size=272, aggression=5, primary_stem="Instrumental")
    net = ref_net(arch, 256, w)
    wave = M.synth_music(20000, seed=41)
    sep = ref_separator(cfg1, net, wave)
    X_ref = sep.loading_mix()
    check("loading_mix 1-band", X_ref, V.loading_mix(wave, cfg1), 1e-4)
    y_ref, v_ref = sep.inference_vr(X_ref.copy(), "cpu", sep.aggressiveness)
    pred = lambda b: V.predict_mask(w, cfg1, b)  # noqa: E731
    y_orc, v_orc = V.inference_vr(X_ref.copy(), cfg1, pred, batch_size=2)
    check("inference_vr y_spec", y_ref, y_orc, 1e-4)
    check("inference_vr v_spec", v_ref, v_orc, 1e-4)
    prim_ref, sec_ref = sep.spec_to_wav(y_ref), sep.spec_to_wav(v_ref)
    prim_orc, sec_orc = V.separate_arrays(wave, cfg1, pred, batch_size=2)
    check("separate 1-band primary", prim_ref, prim_orc, 2e-5)
    check("separate 1-band secondary", sec_ref, sec_orc, 2e-5)
    out.update(wave_seed=41, n_samples=20000, X_1band=X_ref.astype(np.complex64), prim_1band=prim_ref.astype(np.float32), sec_1band=sec_ref.astype(np.float32))
    # test-time augmentation (enable_tta): second pass with the padding shifted by roi/2, averaged
    sep.enable_tta = True
    yt_ref, vt_ref = sep.inference_vr(X_ref.copy(), "cpu", sep.aggressiveness)
    sep.enable_tta = False
    yt_orc, vt_orc = V.inference_vr(X_ref.copy(), cfg1, pred, batch_size=2, enable_tta=True)
    check("inference_vr enable_tta", yt_ref, yt_orc, 1e-4)
    out["y_tta"] = yt_ref.astype(np.complex64)
    # enable_post_process: spec_utils.merge_artifacts on crafted masks (runs at the start, in the middle, close together, at the end, too short)
    su = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.spec_utils")
    crafted = []
    for runs, n in (([(0, 90), (200, 330), (340, 450), (700, 760)], 900), ([(100, 170), (300, 600)], 600), ([(10, 40)], 300), ([], 200), ([(0, 500)], 500)):
        mk = (rng.uniform(0.0, 0.15, (2, 9, n))).astype(np.float32)
        for a_, b_ in runs:
            mk[:, :, a_:b_] = rng.uniform(0.3, 0.9, (2, 9, b_ - a_)).astype(np.float32)
        m_ref = su.merge_artifacts(mk.copy(), thres=0.2)
        check(f"merge_artifacts runs={runs}", m_ref, V.merge_artifacts(mk.copy(), thres=0.2), 0.0)
        crafted.append((mk, m_ref))
    out["pp_mask_in"], out["pp_mask_ref"] = crafted[0]
    sep.enable_post_process, sep.post_process_threshold = True, 0.05
    yp_ref, vp_ref = sep.inference_vr(X_ref.copy(), "cpu", sep.aggressiveness)
    sep.enable_post_process = False
    yp_orc, _ = V.inference_vr(X_ref.copy(), cfg1, pred, batch_size=2, post_process_threshold=0.05)
    check("inference_vr enable_post_process (threshold 0.05)", yp_ref, yp_orc, 1e-4)
    out["y_pp"] = yp_ref.astype(np.complex64)
    # non-accompaniment primary stem (aggressiveness flips) and a vocal-ish aggression
    cfg1v = V.VRConfig(param=p1, nn_architecture=arch, window_size=272, aggression=10, primary_stem="Vocals")
    sepv = ref_separator(cfg1v, net, wave)
    yv_ref, vv_ref = sepv.inference_vr(X_ref.copy(), "cpu", sepv.aggressiveness)
    yv_orc, vv_orc = V.inference_vr(X_ref.copy(), cfg1v, pred, batch_size=3)
    check("inference_vr (Vocals primary, aggression 10)", yv_ref, yv_orc, 1e-4)
    out["y_vocals"] = yv_ref.astype(np.complex64)
    # ---- multi-band: the analysis side and the glue are pinned; the synthesis up-sampling runs the stand-in on both sides
    p4 = V.four_band_v2_param()
    cfg4 = V.VRConfig(param=p4, nn_architecture=33966, window_size=272)
    w4 = V.make_weights(33966, seed=5)
    net4 = ref_net(33966, 1344, w4)
    wave4 = M.synth_music(60000, seed=43)
    sep4 = ref_separator(cfg4, net4, wave4)
    X4_ref = sep4.loading_mix()
    check("loading_mix 4band_v2 (polyphase decimation via scipy)", X4_ref, V.loading_mix(wave4, cfg4), 1e-4)
    y4, v4 = sep4.inference_vr(X4_ref.copy(), "cpu", sep4.aggressiveness)
    pred4 = lambda b: V.predict_mask(w4, cfg4, b)  # noqa: E731
    y4o, v4o = V.inference_vr(X4_ref.copy(), cfg4, pred4, batch_size=2)
    check("inference_vr 4band y_spec", y4, y4o, 1e-4)
    prim4_ref = sep4.spec_to_wav(y4)
    prim4_orc = V.cmb_spectrogram_to_wave(y4o, p4)
    check("cmb_spectrogram_to_wave 4band (glue; up-sampler = stand-in on both sides)", prim4_ref, prim4_orc, 5e-5)
    # high_end_process: the top band's bins above its crop are kept at load time and mirrored back in at synthesis time (vr_separator.py:287-289, :368-372)
    sep4.high_end_process = True
    X4h = sep4.loading_mix()
    hh, he = V.high_end_of(wave4, cfg4)
    assert sep4.input_high_end_h == hh
    check("loading_mix input_high_end", sep4.input_high_end, he[:, :, : sep4.input_high_end.shape[2]], 1e-4)
    prim4h_ref = sep4.spec_to_wav(y4)
    he_m = V.mirroring(y4o, he[:, :, : y4o.shape[2]], p4)
    prim4h_orc = V.cmb_spectrogram_to_wave(y4o, p4, extra_bins_h=hh, extra_bins=he_m)
    check("spec_to_wav high_end_process (mirroring)", prim4h_ref, prim4h_orc, 5e-5)
    sep4.high_end_process, sep4.input_high_end, sep4.input_high_end_h = False, None, None
    out.update(y_4band=np.asarray(y4o).astype(np.complex64), prim_4band_high_end_standin=prim4h_ref.astype(np.float32))
    out.update(wave4_seed=43, n_samples4=60000, X_4band=X4_ref.astype(np.complex64), prim_4band_standin=prim4_ref.astype(np.float32))
    # ---- VR 5.1: CascadedNet (LSTM branch) + the is_v51_model glue (filter masks, convert_channels)
    nets_new = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.vr_network.nets_new")
    for bins, nout, nl in ((128, 8, 16), (64, 16, 32)):
        w51 = V.make_weights_51(bins * 2, nout, nl, seed=3)
        n51 = nets_new.CascadedNet(bins * 2, 56817, nout=nout, nout_lstm=nl).eval()
        assert list(n51.state_dict()) == [n for n, _ in V.param_shapes_51(bins * 2, nout, nl)]
        n51.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w51.items()})
        x51 = np.abs(rng.standard_normal((2, 2, bins + 1, 160))).astype(np.float32)
        with torch.no_grad():
            m51 = n51.predict_mask(torch.from_numpy(x51)).numpy()
        check(f"CascadedNet.predict_mask (VR 5.1, bins {bins}, nout {nout})", m51, V.predict_mask_51(w51, bins * 2, x51), 5e-6)
        if bins == 128:
            out["mask51_in"], out["mask51_ref"] = x51, m51
    p51 = V.single_band_param(n_fft=256, hl=64, bins=128, pre_filter_start=120, pre_filter_stop=127)
    p51["band"][1]["convert_channels"] = "mid_side_c"
    cfg51 = V.VRConfig(param=p51, window_size=160, aggression=5, primary_stem="Vocals", offset=64, is_51=True, nout=8, nout_lstm=16)
    w51 = V.make_weights_51(256, 8, 16, seed=3)
    n51 = nets_new.CascadedNet(256, 56817, nout=8, nout_lstm=16).eval()
    n51.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w51.items()})
    sep51 = ref_separator(cfg51, n51, wave)
    sep51.is_vr_51_model = True
    X51 = sep51.loading_mix()
    check("loading_mix VR 5.1 (convert_channels mid_side_c, lp filter mask)", X51, V.loading_mix(wave, cfg51), 1e-4)
    y51, v51 = sep51.inference_vr(X51.copy(), "cpu", sep51.aggressiveness)
    pred51 = lambda b: V.predict_mask_51(w51, 256, b)  # noqa: E731
    y51o, v51o = V.inference_vr(X51.copy(), cfg51, pred51, batch_size=2)
    check("inference_vr VR 5.1", y51, y51o, 1e-4)
    prim51_ref, sec51_ref = sep51.spec_to_wav(y51), sep51.spec_to_wav(v51)
    prim51, sec51 = V.separate_arrays(wave, cfg51, pred51, batch_size=2)
    check("separate VR 5.1 primary", prim51_ref, prim51, 2e-5)
    check("separate VR 5.1 secondary", sec51_ref, sec51, 2e-5)
    out.update(X_51=X51.astype(np.complex64), prim_51=prim51_ref.astype(np.float32), sec_51=sec51_ref.astype(np.float32))
    # multi-band 5.1 glue (hp / lp filter masks on the synthesis side); the up-sampler is the stand-in on both sides
    p4 = V.four_band_v2_param()
    p4["band"][4]["convert_channels"] = "stereo_n"
    cfg451 = V.VRConfig(param=p4, window_size=160, offset=64, is_51=True)
    sep451 = ref_separator(cfg451, None, wave4)
    sep451.is_vr_51_model = True
    X451 = sep451.loading_mix()
    check("loading_mix 4band VR 5.1", X451, V.loading_mix(wave4, cfg451), 1e-4)
    w451_ref = sep451.spec_to_wav(X451.copy())
    check("cmb_spectrogram_to_wave 4band VR 5.1 (filter masks; stand-in up-sampler)", w451_ref, V.cmb_spectrogram_to_wave(X451.copy(), p4, is_51=True), 5e-5)
    out.update(X_4band_51=X451.astype(np.complex64), wave_4band_51_standin=w451_ref.astype(np.float32))
    np.savez_compressed(os.path.join(GOLD, "vr_small.npz"), **out)
    print("wrote tests/golden/vr_small.npz; oracle pinned: OK (multi-band synthesis up-sampling: parity unpinned)")


if __name__ == "__main__":
    main()
