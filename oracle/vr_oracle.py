"""CPU oracle for the VR (CascadedASPPNet, "VR arch" v4/v5.0) hot path -- TEST INFRASTRUCTURE, not product code.

Restates, with numpy / torch-CPU functional ops:
  * CascadedASPPNet.forward / predict_mask   uvr_lib_v5/vr_network/nets.py:96-175 (+ BaseASPPNet :8-64, capacities :67-93)
  * Conv2DBNActiv / SeperableConv2DBNActiv / Encoder / Decoder / ASPPModule   vr_network/layers.py:8-294 (eval mode: BatchNorm uses
    its running statistics, Dropout2d is the identity)
  * VRSeparator.loading_mix / inference_vr / spec_to_wav   architectures/vr_separator.py:255-375
  * spec_utils.wave_to_spectrogram :282-312, combine_spectrograms :250-279, preprocess :74, make_padding :85, adjust_aggr :472-492,
    cmb_spectrogram_to_wave :341-395, spectrogram_to_wave :315-338, fft_lp_filter / fft_hp_filter :410-429, crop_center :50-71
Third-party pieces the reference calls and this file restates from their published definitions (librosa 0.11 is NOT installed
here):
  * librosa.stft / istft (hann periodic window, center=True, pad_mode="constant", window-sum-square normalisation)
    -- cross-checked against torch.stft / torch.istft by oracle/make_golden_vr.py
  * librosa.resample(res_type="polyphase") == scipy.signal.resample_poly (scipy IS installed: called directly)
  * librosa.resample(res_type="sinc_fastest") (libsamplerate) used by cmb_spectrogram_to_wave for the band up-sampling is NOT
    available: `upsample` below is a Kaiser polyphase stand-in -- PARITY UNPINNED for that one step (multi-band models only).
The network, the inference glue and the single-band path are pinned against the unmodified reference by oracle/make_golden_vr.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SP_ARCH, HP_ARCH, HP2_ARCH = (31191, 33966, 129605), (123821, 123812), (537238, 537227)
NON_ACCOM_STEMS = ("Vocals", "Other", "Bass", "Drums", "Guitar", "Piano", "Synthesizer", "Strings", "Woodwinds", "Brass", "Wind Inst")  # common_separator.py:61


def capacity(nn_architecture: int):
    """determine_model_capacity (nets.py:67-93): (nin, ch) of the four BaseASPPNets; bridges / outputs follow from them."""
    if nn_architecture in SP_ARCH:
        return 16, 8, 16, 32
    if nn_architecture in HP_ARCH:
        return 32, 16, 32, 64
    if nn_architecture in HP2_ARCH:
        return 64, 32, 64, 128
    raise ValueError(f"nn_architecture {nn_architecture} is not a CascadedASPPNet size")


def aspp_extra(nn_architecture: int) -> int:
    """number of extra separable dilated branches (layers.py:240-262): conv6 (and conv7, the SAME module) reuse one `extra_conv`."""
    return 1 if nn_architecture == 129605 else (2 if nn_architecture in (537238, 537227, 33966) else 0)


@dataclass
class VRConfig:
    param: dict  # the modelparams/*.json content (ModelParameters.param, integer band keys)
    nn_architecture: int = 123821
    window_size: int = 512
    aggression: int = 5  # arch_config["aggression"]; value = aggression / 100 (vr_separator.py:67)
    primary_stem: str = "Instrumental"
    offset: int = 128  # CascadedASPPNet.offset (CascadedNet of VR 5.1: 64)
    is_51: bool = False  # VR 5.1 model (model_data has "nout" and "nout_lstm", vr_separator.py:38-41): CascadedNet + the is_v51_model glue
    nout: int = 32
    nout_lstm: int = 128

    @property
    def n_fft_bins(self):  # what VRSeparator passes as n_fft: param["bins"] * 2 (vr_separator.py:178)
        return self.param["bins"] * 2

    @property
    def bands(self):
        return len(self.param["band"])


def single_band_param(n_fft=2048, hl=512, bins=1024, sr=44100, crop_stop=None, pre_filter_start=None, pre_filter_stop=None):
    """modelparams/1band_sr44100_hl512.json by default."""
    return {"bins": bins, "unstable_bins": 0, "reduction_bins": 0, "sr": sr,
            "band": {1: {"sr": sr, "hl": hl, "n_fft": n_fft, "crop_start": 0, "crop_stop": crop_stop or bins, "hpf_start": -1, "res_type": "sinc_best"}},
            "pre_filter_start": bins - 1 if pre_filter_start is None else pre_filter_start, "pre_filter_stop": bins if pre_filter_stop is None else pre_filter_stop,
            "mid_side": False, "mid_side_b": False, "mid_side_b2": False, "stereo_w": False, "stereo_n": False, "reverse": False}


def four_band_v2_param():
    """modelparams/4band_v2.json."""
    return {"bins": 672, "unstable_bins": 8, "reduction_bins": 637, "sr": 44100, "pre_filter_start": 668, "pre_filter_stop": 672,
            "band": {1: {"sr": 7350, "hl": 80, "n_fft": 640, "crop_start": 0, "crop_stop": 85, "lpf_start": 25, "lpf_stop": 53, "res_type": "polyphase"},
                     2: {"sr": 7350, "hl": 80, "n_fft": 320, "crop_start": 4, "crop_stop": 87, "hpf_start": 25, "hpf_stop": 12, "lpf_start": 31, "lpf_stop": 62, "res_type": "polyphase"},
                     3: {"sr": 14700, "hl": 160, "n_fft": 512, "crop_start": 17, "crop_stop": 216, "hpf_start": 48, "hpf_stop": 24, "lpf_start": 139, "lpf_stop": 210, "res_type": "polyphase"},
                     4: {"sr": 44100, "hl": 480, "n_fft": 960, "crop_start": 78, "crop_stop": 383, "hpf_start": 130, "hpf_stop": 86, "res_type": "kaiser_fast"}},
            "mid_side": False, "mid_side_b": False, "mid_side_b2": False, "stereo_w": False, "stereo_n": False, "reverse": False}


# --------------------------------------------------------------------------------------------------------- the network
def param_shapes(nn_architecture: int):
    """(name, shape) in the reference module's state_dict order (BatchNorm buffers included)."""
    out = []

    def cba(p, nin, nout, k):  # Conv2DBNActiv
        out.append((f"{p}.conv.0.weight", (nout, nin, k, k)))
        for nm in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{p}.conv.1.{nm}", (nout,)))
        out.append((f"{p}.conv.1.num_batches_tracked", ()))

    def sep(p, nin, nout):  # SeperableConv2DBNActiv
        out.append((f"{p}.conv.0.weight", (nin, 1, 3, 3)))
        out.append((f"{p}.conv.1.weight", (nout, nin, 1, 1)))
        for nm in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{p}.conv.2.{nm}", (nout,)))
        out.append((f"{p}.conv.2.num_batches_tracked", ()))

    def aspp(p, nin, nout):
        cba(f"{p}.conv1.1", nin, nin, 1)
        cba(f"{p}.conv2", nin, nin, 1)
        for i in (3, 4, 5):
            sep(f"{p}.conv{i}", nin, nin)
        for i in range(aspp_extra(nn_architecture)):
            sep(f"{p}.conv{6 + i}", nin, nin)
        cba(f"{p}.bottleneck.0", nin * (5 + aspp_extra(nn_architecture)), nout, 1)

    def base(p, nin, ch):
        c = nin
        for i, m in enumerate((1, 2, 4, 8)):
            cba(f"{p}.enc{i + 1}.conv1", c, ch * m, 3)
            cba(f"{p}.enc{i + 1}.conv2", ch * m, ch * m, 3)
            c = ch * m
        if nn_architecture == 129605:
            cba(f"{p}.enc5.conv1", ch * 8, ch * 16, 3)
            cba(f"{p}.enc5.conv2", ch * 16, ch * 16, 3)
            aspp(f"{p}.aspp", ch * 16, ch * 32)
            cba(f"{p}.dec5.conv", ch * 48, ch * 16, 3)
        else:
            aspp(f"{p}.aspp", ch * 8, ch * 16)
        for i, (a, b) in zip((4, 3, 2, 1), ((24, 8), (12, 4), (6, 2), (3, 1))):
            cba(f"{p}.dec{i}.conv", ch * a, ch * b, 3)

    c1, cb, c2, c3 = capacity(nn_architecture)
    base("stg1_low_band_net", 2, c1)
    base("stg1_high_band_net", 2, c1)
    cba("stg2_bridge", c1 + 2, cb, 1)
    base("stg2_full_band_net", cb, c2)
    cba("stg3_bridge", c1 + c2 + 2, c2, 1)
    base("stg3_full_band_net", c2, c3)
    out.append(("out.weight", (2, c3, 1, 1)))
    out.append(("aux1_out.weight", (2, c1, 1, 1)))
    out.append(("aux2_out.weight", (2, c2, 1, 1)))
    return out


def make_weights(nn_architecture: int, seed=0):
    rng = np.random.default_rng(seed)
    w = {}
    shared = {}
    for name, shape in param_shapes(nn_architecture):
        if name.endswith("num_batches_tracked"):
            a = np.array(100, dtype=np.int64)
        elif name.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif name.endswith("running_mean"):
            a = rng.normal(0.0, 0.2, shape).astype(np.float32)
        elif len(shape) == 1 and name.endswith(".weight"):
            a = rng.uniform(0.7, 1.3, shape).astype(np.float32)
        elif len(shape) == 1:
            a = rng.normal(0.0, 0.1, shape).astype(np.float32)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            a = rng.normal(0.0, math.sqrt(2.0 / fan_in), shape).astype(np.float32)
        if ".aspp.conv7." in name:  # conv7 IS conv6 in the reference module (one `extra_conv` registered twice)
            a = shared[name.replace(".aspp.conv7.", ".aspp.conv6.")]
        if ".aspp.conv6." in name:
            shared[name] = a
        w[name] = a
    return w


def net_forward(weights, nn_architecture: int, n_fft_bins: int, x: np.ndarray, dtype="float32") -> np.ndarray:
    """CascadedASPPNet.forward in eval mode: x (B, 2, bins + 1, W) magnitudes -> mask (B, 2, bins + 1, W)."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items() if not k.endswith("num_batches_tracked")}
    x = torch.from_numpy(np.ascontiguousarray(x)).to(td)

    def bn_act(y, p, leaky):
        y = F.batch_norm(y, W[f"{p}.running_mean"], W[f"{p}.running_var"], W[f"{p}.weight"], W[f"{p}.bias"], False, 0.0, 1e-5)
        return F.leaky_relu(y, 0.01) if leaky else F.relu(y)

    def cba(y, p, stride=1, pad=1, leaky=False):
        k = W[f"{p}.conv.0.weight"].shape[-1]
        return bn_act(F.conv2d(y, W[f"{p}.conv.0.weight"], stride=stride, padding=pad if k == 3 else 0), f"{p}.conv.1", leaky)

    def sep(y, p, dil):
        y = F.conv2d(y, W[f"{p}.conv.0.weight"], padding=dil, dilation=dil, groups=y.shape[1])
        return bn_act(F.conv2d(y, W[f"{p}.conv.1.weight"]), f"{p}.conv.2", False)

    def aspp(y, p):
        h, w_ = y.shape[2:]
        f1 = F.interpolate(cba(F.adaptive_avg_pool2d(y, (1, None)), f"{p}.conv1.1"), size=(h, w_), mode="bilinear", align_corners=True)
        feats = [f1, cba(y, f"{p}.conv2"), sep(y, f"{p}.conv3", 4), sep(y, f"{p}.conv4", 8), sep(y, f"{p}.conv5", 16)]
        for i in range(aspp_extra(nn_architecture)):
            feats.append(sep(y, f"{p}.conv{6 + i}", 16))
        return cba(torch.cat(feats, 1), f"{p}.bottleneck.0")

    def dec(y, skip, p):
        y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
        d = skip.shape[3] - y.shape[3]  # crop_center: time axis only (spec_utils.py:50-71)
        assert d >= 0
        if d:
            skip = skip[:, :, :, d // 2 : d // 2 + y.shape[3]]
        return cba(torch.cat([y, skip], 1), f"{p}.conv")

    def base(y, p):
        skips = []
        n_enc = 5 if nn_architecture == 129605 else 4
        for i in range(1, n_enc + 1):
            s = cba(y, f"{p}.enc{i}.conv1", leaky=True)
            y = cba(s, f"{p}.enc{i}.conv2", stride=2, leaky=True)
            skips.append(s)
        y = aspp(y, f"{p}.aspp")
        for i in range(n_enc, 0, -1):
            y = dec(y, skips[i - 1], f"{p}.dec{i}")
        return y

    with torch.no_grad():
        max_bin, output_bin = n_fft_bins // 2, n_fft_bins // 2 + 1
        x = x[:, :, :max_bin]
        bw = x.shape[2] // 2
        aux1 = torch.cat([base(x[:, :, :bw], "stg1_low_band_net"), base(x[:, :, bw:], "stg1_high_band_net")], 2)
        aux2 = base(cba(torch.cat([x, aux1], 1), "stg2_bridge"), "stg2_full_band_net")
        h = base(cba(torch.cat([x, aux1, aux2], 1), "stg3_bridge"), "stg3_full_band_net")
        mask = torch.sigmoid(F.conv2d(h, W["out.weight"]))
        mask = F.pad(mask, (0, 0, 0, output_bin - mask.shape[2]), mode="replicate")
    return mask.to(torch.float32).numpy()


def predict_mask(weights, cfg: VRConfig, x: np.ndarray, dtype="float32") -> np.ndarray:
    m = net_forward(weights, cfg.nn_architecture, cfg.n_fft_bins, x, dtype)
    return m[:, :, :, cfg.offset : -cfg.offset] if cfg.offset > 0 else m


# --------------------------------------------------------------------------------------------------------- librosa restatements
def hann_periodic(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def stft(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """librosa.stft(y, n_fft=, hop_length=) defaults: hann (periodic), center=True, pad_mode="constant", complex64.  (n,) -> (n_fft/2+1, 1 + n//hop)"""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, (n_fft // 2, n_fft // 2))
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx] * hann_periodic(n_fft)[:, None]
    return np.fft.rfft(frames.astype(np.float32), axis=0).astype(np.complex64)


def istft(spec: np.ndarray, hop: int, length=None) -> np.ndarray:
    """librosa.istft(spec, hop_length=, length=) defaults: n_fft from the bin count, hann, center=True; length=None -> hop*(frames-1) samples,
    otherwise the overlap-add buffer is read from n_fft/2 for `length` samples (zero-padded past its end)."""
    n_fft = 2 * (spec.shape[0] - 1)
    n_frames = spec.shape[1]
    win = hann_periodic(n_fft)
    frames = np.fft.irfft(spec, n=n_fft, axis=0).astype(np.float32) * win[:, None]
    total = n_fft + hop * (n_frames - 1)
    y = np.zeros(total, np.float32)
    wss = np.zeros(total, np.float32)
    for t in range(n_frames):
        y[t * hop : t * hop + n_fft] += frames[:, t]
        wss[t * hop : t * hop + n_fft] += win * win
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    if length is not None:
        out = y[n_fft // 2 : n_fft // 2 + length]
        return np.pad(out, (0, length - len(out))) if len(out) < length else out
    return y[n_fft // 2 : total - n_fft // 2]


def resample_polyphase(y: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """librosa.resample(res_type="polyphase"): scipy.signal.resample_poly with the gcd-reduced integer ratio, float32 out."""
    if orig_sr == target_sr:
        return y
    import scipy.signal

    g = math.gcd(int(orig_sr), int(target_sr))
    return scipy.signal.resample_poly(y, int(target_sr) // g, int(orig_sr) // g, axis=-1).astype(np.float32)


def upsample(y: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """STAND-IN for librosa.resample(res_type="sinc_fastest") (libsamplerate, absent): Kaiser polyphase.  PARITY UNPINNED."""
    return resample_polyphase(np.asarray(y, dtype=np.float32), orig_sr, target_sr)


# --------------------------------------------------------------------------------------------------------- spec_utils / VRSeparator
def convert_channels(spec, param, band):  # spec_utils.py:232-247
    cc = param["band"][band].get("convert_channels")
    if cc == "mid_side_c":
        return np.stack([spec[0] + spec[1] * 0.25, spec[1] - spec[0] * 0.25])
    if cc == "mid_side":
        return np.stack([(spec[0] + spec[1]) / 2, spec[0] - spec[1]])
    if cc == "stereo_n":
        return np.stack([(spec[0] + spec[1] * 0.25) / 0.9375, (spec[1] + spec[0] * 0.25) / 0.9375])
    return spec


def wave_to_spectrogram(wave: np.ndarray, hop: int, n_fft: int, param: dict, is_51=False, band=None) -> np.ndarray:
    """spec_utils.py:282-312."""
    if is_51:
        return convert_channels(np.stack([stft(wave[0], n_fft, hop), stft(wave[1], n_fft, hop)]), param, band)
    if param.get("reverse"):
        l, r = np.flip(wave[0]), np.flip(wave[1])
    elif param.get("mid_side"):
        l, r = (wave[0] + wave[1]) / 2, wave[0] - wave[1]
    elif param.get("mid_side_b2"):
        l, r = wave[1] + wave[0] * 0.5, wave[0] - wave[1] * 0.5
    else:
        l, r = wave[0], wave[1]
    return np.stack([stft(l, n_fft, hop), stft(r, n_fft, hop)])


def loading_mix(wave: np.ndarray, cfg: VRConfig) -> np.ndarray:
    """VRSeparator.loading_mix (vr_separator.py:255-293) for an input already at the top band's rate: (2, N) -> (2, bins+1, frames) c64."""
    p = cfg.param
    n = cfg.bands
    assert p["band"][n]["sr"] == p["sr"], "the top band is read at the file's rate"
    waves, specs = {}, {}
    for d in range(n, 0, -1):
        bp = p["band"][d]
        if d == n:
            waves[d] = np.asarray(wave, dtype=np.float32)
        else:
            if bp["res_type"] != "polyphase" and bp["sr"] != p["band"][d + 1]["sr"]:
                raise NotImplementedError(f"band {d} resamples with res_type={bp['res_type']}: only polyphase is restated")
            waves[d] = resample_polyphase(waves[d + 1], p["band"][d + 1]["sr"], bp["sr"])
        specs[d] = wave_to_spectrogram(waves[d], bp["hl"], bp["n_fft"], p, cfg.is_51, d)
    return combine_spectrograms(specs, p, cfg.is_51)


def combine_spectrograms(specs: dict, p: dict, is_51=False) -> np.ndarray:
    """spec_utils.py:250-279."""
    n = len(p["band"])
    l = min(specs[i].shape[2] for i in specs)
    out = np.zeros((2, p["bins"] + 1, l), np.complex64)
    off = 0
    for d in range(1, n + 1):
        bp = p["band"][d]
        h = bp["crop_stop"] - bp["crop_start"]
        out[:, off : off + h, :] = specs[d][:, bp["crop_start"] : bp["crop_stop"], :l]
        off += h
    if off > p["bins"]:
        raise ValueError("Too much bins")
    if p["pre_filter_start"] > 0:
        if is_51:
            out = (out * lp_filter_mask(out.shape[1], p["pre_filter_start"], p["pre_filter_stop"])).astype(np.complex64)
        elif n == 1:
            out = fft_lp_filter(out, p["pre_filter_start"], p["pre_filter_stop"])
        else:
            gp = 1.0
            for b in range(p["pre_filter_start"] + 1, p["pre_filter_stop"]):
                g = math.pow(10, -(b - p["pre_filter_start"]) * (3.5 - gp) / 20.0)
                gp = g
                out[:, b, :] *= g
    return out


def fft_lp_filter(spec, bin_start, bin_stop):  # spec_utils.py:410-418
    g = 1.0
    for b in range(bin_start, bin_stop):
        g -= 1 / (bin_stop - bin_start)
        spec[:, b, :] = g * spec[:, b, :]
    spec[:, bin_stop:, :] *= 0
    return spec


def fft_hp_filter(spec, bin_start, bin_stop):  # spec_utils.py:421-429
    g = 1.0
    for b in range(bin_start, bin_stop, -1):
        g -= 1 / (bin_start - bin_stop)
        spec[:, b, :] = g * spec[:, b, :]
    spec[:, 0 : bin_stop + 1, :] *= 0
    return spec


def make_padding(width, cropsize, offset):  # spec_utils.py:85-96
    roi = cropsize - offset * 2
    if roi == 0:
        roi = cropsize
    return offset, roi - (width % roi) + offset, roi


def adjust_aggr(mask, is_non_accom_stem, value, split_bin, aggr_correction=None):  # spec_utils.py:472-492
    aggr = value * 2
    if aggr != 0:
        if is_non_accom_stem:
            aggr = 1 - aggr
        a = [aggr, aggr]
        if aggr_correction is not None:
            a[0] += aggr_correction["left"]
            a[1] += aggr_correction["right"]
        for ch in range(2):
            mask[ch, :split_bin] = np.power(mask[ch, :split_bin], 1 + a[ch] / 3)
            mask[ch, split_bin:] = np.power(mask[ch, split_bin:], 1 + a[ch])
    return mask


def merge_artifacts(y_mask, thres=0.01, min_range=64, fade_size=32):
    """spec_utils.merge_artifacts (:180-223): frames whose smallest mask value exceeds `thres` for more than `min_range` consecutive frames are pushed
    towards 1 with linear fades at the run ends (enable_post_process).  Any exception inside leaves the mask unchanged, as in the reference."""
    mask = y_mask
    try:
        if min_range < fade_size * 2:
            raise ValueError("min_range must be >= fade_size * 2")
        idx = np.where(y_mask.min(axis=(0, 1)) > thres)[0]
        start_idx = np.insert(idx[np.where(np.diff(idx) != 1)[0] + 1], 0, idx[0])
        end_idx = np.append(idx[np.where(np.diff(idx) != 1)[0]], idx[-1])
        artifact_idx = np.where(end_idx - start_idx > min_range)[0]
        weight = np.zeros_like(y_mask)
        if len(artifact_idx) > 0:
            start_idx = start_idx[artifact_idx]
            end_idx = end_idx[artifact_idx]
            old_e = None
            for s, e in zip(start_idx, end_idx):
                if old_e is not None and s - old_e < fade_size:
                    s = old_e - fade_size * 2
                if s != 0:
                    weight[:, :, s : s + fade_size] = np.linspace(0, 1, fade_size)
                else:
                    s -= fade_size
                if e != y_mask.shape[2]:
                    weight[:, :, e - fade_size : e] = np.linspace(1, 0, fade_size)
                else:
                    e += fade_size
                weight[:, :, s + fade_size : e - fade_size] = 1
                old_e = e
        v_mask = 1 - y_mask
        y_mask += weight * v_mask
        mask = y_mask
    except Exception:  # noqa: BLE001  (the reference prints and carries on)
        pass
    return mask


def inference_vr(X_spec: np.ndarray, cfg: VRConfig, predict, batch_size=1, enable_tta=False, post_process_threshold=None):
    """VRSeparator.inference_vr (vr_separator.py:295-366; enable_post_process off): -> (y_spec, v_spec) complex.
    predict: (B, 2, bins+1, window) float32 -> (B, 2, bins+1, window - 2*offset)."""
    X_mag, X_phase = np.abs(X_spec), np.angle(X_spec)
    n_frame = X_mag.shape[2]
    pad_l, pad_r, roi = make_padding(n_frame, cfg.window_size, cfg.offset)

    def execute(pl, pr):  # _execute (:296-327)
        X_pad = np.pad(X_mag, ((0, 0), (0, 0), (pl, pr)), mode="constant")
        X_pad /= X_pad.max()
        patches = (X_pad.shape[2] - 2 * cfg.offset) // roi
        data = np.asarray([X_pad[:, :, i * roi : i * roi + cfg.window_size] for i in range(patches)])
        masks = []
        for i in range(0, patches, batch_size):
            pred = predict(data[i : i + batch_size])
            masks.append(np.concatenate(list(pred), axis=2))
        return np.concatenate(masks, axis=2)

    mask = execute(pad_l, pad_r)
    if enable_tta:  # test-time augmentation: a second pass shifted by half a region of interest, averaged (:351-359)
        mask_tta = execute(pad_l + roi // 2, pad_r + roi // 2)[:, :, roi // 2 :]
        mask = (mask[:, :, :n_frame] + mask_tta[:, :, :n_frame]) * 0.5
    else:
        mask = mask[:, :, :n_frame]
    value = float(int(cfg.aggression) / 100)
    mask = adjust_aggr(mask, cfg.primary_stem in NON_ACCOM_STEMS, value, cfg.param["band"][1]["crop_stop"], cfg.param.get("aggr_correction"))
    if post_process_threshold is not None:  # enable_post_process (:334-335)
        mask = merge_artifacts(mask, thres=post_process_threshold)
    y = mask * X_mag * np.exp(1.0j * X_phase)
    v = (1 - mask) * X_mag * np.exp(1.0j * X_phase)
    return y, v


def spectrogram_to_wave(spec, hop, param, is_51=False, band=None):  # spec_utils.py:315-338
    l, r = istft(spec[0], hop), istft(spec[1], hop)
    if is_51:
        cc = param["band"][band].get("convert_channels")
        if cc == "mid_side_c":
            return np.stack([l / 1.0625 - r / 4.25, r / 1.0625 + l / 4.25])
        if cc == "mid_side":
            return np.stack([l + r / 2, l - r / 2])
        if cc == "stereo_n":
            return np.stack([l - r * 0.25, r - l * 0.25])
        return np.stack([l, r])
    if param.get("reverse"):
        return np.stack([np.flip(l), np.flip(r)])
    if param.get("mid_side"):
        return np.stack([l + r / 2, l - r / 2])
    if param.get("mid_side_b2"):
        return np.stack([r / 1.25 + 0.4 * l, l / 1.25 - 0.4 * r])
    return np.stack([l, r])


def high_end_of(wave: np.ndarray, cfg: "VRConfig"):
    """The slice loading_mix keeps when high_end_process is on (vr_separator.py:287-289): the top band's own STFT bins above its crop."""
    p = cfg.param
    n = cfg.bands
    bp = p["band"][n]
    h = (bp["n_fft"] // 2 - bp["crop_stop"]) + (p["pre_filter_stop"] - p["pre_filter_start"])
    spec = wave_to_spectrogram(np.asarray(wave, dtype=np.float32), bp["hl"], bp["n_fft"], p, cfg.is_51, n)
    return h, spec[:, bp["n_fft"] // 2 - h : bp["n_fft"] // 2, :]


def mirroring(spec_m, input_high_end, p):
    """spec_utils.mirroring("mirroring", ...) (:458-463): the magnitudes just below the pre-filter, flipped, with the input's phases, where they are smaller than the input."""
    pfs = p["pre_filter_start"]
    mirror = np.flip(np.abs(spec_m[:, pfs - 10 - input_high_end.shape[1] : pfs - 10, :]), 1)
    mirror = mirror * np.exp(1.0j * np.angle(input_high_end))
    return np.where(np.abs(input_high_end) <= np.abs(mirror), input_high_end, mirror)


def cmb_spectrogram_to_wave(spec_m: np.ndarray, p: dict, up=upsample, is_51=False, extra_bins_h=None, extra_bins=None) -> np.ndarray:
    """spec_utils.py:341-395.  `up` is the band up-sampler (see the module docstring)."""
    def hp(s_, a, b_):
        return s_ * hp_filter_mask(s_.shape[1], a, b_) if is_51 else fft_hp_filter(s_, a, b_)

    def lp(s_, a, b_):
        return s_ * lp_filter_mask(s_.shape[1], a, b_) if is_51 else fft_lp_filter(s_, a, b_)

    n = len(p["band"])
    off = 0
    wave = None
    for d in range(1, n + 1):
        bp = p["band"][d]
        s = np.zeros((2, bp["n_fft"] // 2 + 1, spec_m.shape[2]), dtype=complex)
        h = bp["crop_stop"] - bp["crop_start"]
        s[:, bp["crop_start"] : bp["crop_stop"], :] = spec_m[:, off : off + h, :]
        off += h
        if d == n:
            if extra_bins_h:  # high_end_process (spec_utils.py:354-356)
                max_bin = bp["n_fft"] // 2
                s[:, max_bin - extra_bins_h : max_bin, :] = extra_bins[:, :extra_bins_h, :]
            if bp["hpf_start"] > 0:
                s = hp(s, bp["hpf_start"], bp["hpf_stop"] - 1)
            w_d = spectrogram_to_wave(s, bp["hl"], p, is_51, d)
            wave = w_d if n == 1 else np.add(wave, w_d)
        else:
            sr = p["band"][d + 1]["sr"]
            if d == 1:
                s = lp(s, bp["lpf_start"], bp["lpf_stop"])
                wave = up(spectrogram_to_wave(s, bp["hl"], p, is_51, d), bp["sr"], sr)
            else:
                s = hp(s, bp["hpf_start"], bp["hpf_stop"] - 1)
                s = lp(s, bp["lpf_start"], bp["lpf_stop"])
                wave = up(np.add(wave, spectrogram_to_wave(s, bp["hl"], p, is_51, d)), bp["sr"], sr)
    return wave


def separate_arrays(wave: np.ndarray, cfg: VRConfig, predict, batch_size=1, up=upsample):
    """VRSeparator.separate between reading the file and final_process (vr_separator.py:183-224): (2, N) -> primary (2, M), secondary (2, M)."""
    X = loading_mix(wave, cfg)
    y, v = inference_vr(X, cfg, predict, batch_size)
    y = np.nan_to_num(y, nan=0.0, posinf=0.0, neginf=0.0)
    v = np.nan_to_num(v, nan=0.0, posinf=0.0, neginf=0.0)
    return cmb_spectrogram_to_wave(y, cfg.param, up, cfg.is_51).astype(np.float32), cmb_spectrogram_to_wave(v, cfg.param, up, cfg.is_51).astype(np.float32)


# =========================================================================================================================
# VR 5.1: CascadedNet with an LSTM branch (uvr_lib_v5/vr_network/nets_new.py:8-160, layers_new.py:8-149) and the is_v51_model variants of the
# spectrogram glue (spec_utils.py: convert_channels :232-247, get_lp_filter_mask / get_hp_filter_mask :398-407, the `is_v51_model` branches of
# combine_spectrograms :266-268, wave_to_spectrogram :301-310, spectrogram_to_wave :322-330, cmb_spectrogram_to_wave :357-384).
def param_shapes_51(n_fft_bins: int, nout: int, nout_lstm: int, nn_arch_size: int = 56817):
    out = []
    max_bin = n_fft_bins // 2
    nin_lstm = max_bin // 2
    nout = 64 if nn_arch_size == 218409 else nout

    def cba(p, nin, no, k):
        out.append((f"{p}.conv.0.weight", (no, nin, k, k)))
        for nm in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{p}.conv.1.{nm}", (no,)))
        out.append((f"{p}.conv.1.num_batches_tracked", ()))

    def base(p, nin, no, n_lstm_in, n_lstm_out):
        cba(f"{p}.enc1", nin, no, 3)
        c = no
        for i, m in zip((2, 3, 4, 5), (2, 4, 6, 8)):
            cba(f"{p}.enc{i}.conv1", c, no * m, 3)
            cba(f"{p}.enc{i}.conv2", no * m, no * m, 3)
            c = no * m
        cba(f"{p}.aspp.conv1.1", no * 8, no * 8, 1)
        cba(f"{p}.aspp.conv2", no * 8, no * 8, 1)
        for i in (3, 4, 5):
            cba(f"{p}.aspp.conv{i}", no * 8, no * 8, 3)
        cba(f"{p}.aspp.bottleneck", no * 40, no * 8, 1)
        cba(f"{p}.dec4.conv1", no * 14, no * 6, 3)
        cba(f"{p}.dec3.conv1", no * 10, no * 4, 3)
        cba(f"{p}.dec2.conv1", no * 6, no * 2, 3)
        cba(f"{p}.lstm_dec2.conv", no * 2, 1, 1)
        hid = n_lstm_out // 2
        for sfx in ("", "_reverse"):
            out.extend([(f"{p}.lstm_dec2.lstm.weight_ih_l0{sfx}", (4 * hid, n_lstm_in)), (f"{p}.lstm_dec2.lstm.weight_hh_l0{sfx}", (4 * hid, hid)),
                        (f"{p}.lstm_dec2.lstm.bias_ih_l0{sfx}", (4 * hid,)), (f"{p}.lstm_dec2.lstm.bias_hh_l0{sfx}", (4 * hid,))])
        out.extend([(f"{p}.lstm_dec2.dense.0.weight", (n_lstm_in, n_lstm_out)), (f"{p}.lstm_dec2.dense.0.bias", (n_lstm_in,))])
        for nm in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{p}.lstm_dec2.dense.1.{nm}", (n_lstm_in,)))
        out.append((f"{p}.lstm_dec2.dense.1.num_batches_tracked", ()))
        cba(f"{p}.dec1.conv1", no * 3 + 1, no, 3)

    base("stg1_low_band_net.0", 2, nout // 2, nin_lstm // 2, nout_lstm)
    cba("stg1_low_band_net.1", nout // 2, nout // 4, 1)
    base("stg1_high_band_net", 2, nout // 4, nin_lstm // 2, nout_lstm // 2)
    base("stg2_low_band_net.0", nout // 4 + 2, nout, nin_lstm // 2, nout_lstm)
    cba("stg2_low_band_net.1", nout, nout // 2, 1)
    base("stg2_high_band_net", nout // 4 + 2, nout // 2, nin_lstm // 2, nout_lstm // 2)
    base("stg3_full_band_net", 3 * nout // 4 + 2, nout, nin_lstm, nout_lstm)
    out.append(("out.weight", (2, nout, 1, 1)))
    out.append(("aux_out.weight", (2, 3 * nout // 4, 1, 1)))
    return out


def make_weights_51(n_fft_bins, nout, nout_lstm, seed=0, nn_arch_size=56817):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in param_shapes_51(n_fft_bins, nout, nout_lstm, nn_arch_size):
        if name.endswith("num_batches_tracked"):
            a = np.array(100, dtype=np.int64)
        elif name.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif name.endswith("running_mean"):
            a = rng.normal(0.0, 0.2, shape).astype(np.float32)
        elif ".lstm." in name:
            a = rng.uniform(-0.3, 0.3, shape).astype(np.float32)
        elif len(shape) == 1 and name.endswith(".weight"):
            a = rng.uniform(0.7, 1.3, shape).astype(np.float32)
        elif len(shape) == 1:
            a = rng.normal(0.0, 0.1, shape).astype(np.float32)
        elif len(shape) == 2:
            a = rng.normal(0.0, math.sqrt(1.0 / shape[1]), shape).astype(np.float32)
        else:
            a = rng.normal(0.0, math.sqrt(2.0 / (shape[1] * shape[2] * shape[3])), shape).astype(np.float32)
        w[name] = a
    return w


def net_forward_51(weights, n_fft_bins: int, x: np.ndarray, dtype="float32") -> np.ndarray:
    """CascadedNet.forward in eval mode (nets_new.py:104-137): x (B, 2, bins + 1, W) -> mask (B, 2, bins + 1, W)."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items() if not k.endswith("num_batches_tracked")}
    x = torch.from_numpy(np.ascontiguousarray(x)).to(td)

    def cba(y, p, stride=1, pad=None, dil=1, leaky=False):
        k = W[f"{p}.conv.0.weight"].shape[-1]
        if pad is None:
            pad = 1 if k == 3 else 0
        y = F.conv2d(y, W[f"{p}.conv.0.weight"], stride=stride, padding=pad, dilation=dil)
        y = F.batch_norm(y, W[f"{p}.conv.1.running_mean"], W[f"{p}.conv.1.running_var"], W[f"{p}.conv.1.weight"], W[f"{p}.conv.1.bias"], False, 0.0, 1e-5)
        return F.leaky_relu(y, 0.01) if leaky else F.relu(y)

    def enc(y, p, stride):
        return cba(cba(y, f"{p}.conv1", stride=stride, leaky=True), f"{p}.conv2", leaky=True)

    def dec(y, skip, p):
        y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
        d = skip.shape[3] - y.shape[3]
        assert d >= 0
        if d:
            skip = skip[:, :, :, d // 2 : d // 2 + y.shape[3]]
        return cba(torch.cat([y, skip], 1), f"{p}.conv1")

    def aspp(y, p):
        h, w_ = y.shape[2:]
        f1 = F.interpolate(cba(F.adaptive_avg_pool2d(y, (1, None)), f"{p}.conv1.1"), size=(h, w_), mode="bilinear", align_corners=True)
        feats = [f1, cba(y, f"{p}.conv2")] + [cba(y, f"{p}.conv{i}", pad=dl, dil=dl) for i, dl in zip((3, 4, 5), ((4, 2), (8, 4), (12, 6)))]
        return cba(torch.cat(feats, 1), f"{p}.bottleneck")

    def lstm_dir(seq, p, sfx):  # seq (T, N, in) -> (T, N, hid); gate order i, f, g, o
        Wi, Wh, bi, bh = W[f"{p}.weight_ih_l0{sfx}"], W[f"{p}.weight_hh_l0{sfx}"], W[f"{p}.bias_ih_l0{sfx}"], W[f"{p}.bias_hh_l0{sfx}"]
        hid = Wh.shape[1]
        h = torch.zeros(seq.shape[1], hid, dtype=td)
        c = torch.zeros_like(h)
        outs = []
        for t in range(seq.shape[0]):
            g = seq[t] @ Wi.t() + bi + h @ Wh.t() + bh
            i_, f_, g_, o_ = g.chunk(4, dim=1)
            c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
            h = torch.sigmoid(o_) * torch.tanh(c)
            outs.append(h)
        return torch.stack(outs)

    def lstm_module(y, p):  # layers_new.py:130-149
        N, _, nbins, nframes = y.shape
        h = cba(y, f"{p}.conv")[:, 0].permute(2, 0, 1)  # nframes, N, nbins
        fw = lstm_dir(h, f"{p}.lstm", "")
        bw = lstm_dir(h.flip(0), f"{p}.lstm", "_reverse").flip(0)
        h = torch.cat([fw, bw], dim=-1).reshape(nframes * N, -1)
        h = F.linear(h, W[f"{p}.dense.0.weight"], W[f"{p}.dense.0.bias"])
        h = F.relu(F.batch_norm(h, W[f"{p}.dense.1.running_mean"], W[f"{p}.dense.1.running_var"], W[f"{p}.dense.1.weight"], W[f"{p}.dense.1.bias"], False, 0.0, 1e-5))
        return h.reshape(nframes, N, 1, nbins).permute(1, 2, 3, 0)

    def base(y, p):
        e1 = cba(y, f"{p}.enc1")
        e2 = enc(e1, f"{p}.enc2", 2)
        e3 = enc(e2, f"{p}.enc3", 2)
        e4 = enc(e3, f"{p}.enc4", 2)
        e5 = enc(e4, f"{p}.enc5", 2)
        h = aspp(e5, f"{p}.aspp")
        h = dec(h, e4, f"{p}.dec4")
        h = dec(h, e3, f"{p}.dec3")
        h = dec(h, e2, f"{p}.dec2")
        h = torch.cat([h, lstm_module(h, f"{p}.lstm_dec2")], 1)
        return dec(h, e1, f"{p}.dec1")

    with torch.no_grad():
        max_bin, output_bin = n_fft_bins // 2, n_fft_bins // 2 + 1
        x = x[:, :, :max_bin]
        bw = x.shape[2] // 2
        l1_in, h1_in = x[:, :, :bw], x[:, :, bw:]
        l1 = cba(base(l1_in, "stg1_low_band_net.0"), "stg1_low_band_net.1")
        h1 = base(h1_in, "stg1_high_band_net")
        aux1 = torch.cat([l1, h1], 2)
        l2 = cba(base(torch.cat([l1_in, l1], 1), "stg2_low_band_net.0"), "stg2_low_band_net.1")
        h2 = base(torch.cat([h1_in, h1], 1), "stg2_high_band_net")
        aux2 = torch.cat([l2, h2], 2)
        f3 = base(torch.cat([x, aux1, aux2], 1), "stg3_full_band_net")
        mask = torch.sigmoid(F.conv2d(f3, W["out.weight"]))
        mask = F.pad(mask, (0, 0, 0, output_bin - mask.shape[2]), mode="replicate")
    return mask.to(torch.float32).numpy()


def predict_mask_51(weights, n_fft_bins, x, offset=64, dtype="float32"):
    m = net_forward_51(weights, n_fft_bins, x, dtype)
    return m[:, :, :, offset:-offset] if offset > 0 else m


def lp_filter_mask(n_bins, bin_start, bin_stop):  # spec_utils.py:398-401
    return np.concatenate([np.ones((bin_start - 1, 1)), np.linspace(1, 0, bin_stop - bin_start + 1)[:, None], np.zeros((n_bins - bin_stop, 1))], axis=0)


def hp_filter_mask(n_bins, bin_start, bin_stop):  # spec_utils.py:404-407
    return np.concatenate([np.zeros((bin_stop + 1, 1)), np.linspace(0, 1, 1 + bin_start - bin_stop)[:, None], np.ones((n_bins - bin_start - 2, 1))], axis=0)
