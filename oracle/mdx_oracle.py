"""CPU oracle for the MDX hot path (TEST INFRASTRUCTURE -- not product code).

A plain numpy / torch-CPU restatement of the reference's per-chunk MDX loop:

    frame -> STFT -> ConvTDFNet forward -> iSTFT -> Hann-window overlap-add -> normalise

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  The product path (python-audio-separator_b200/) never imports it.

Pinning: oracle/make_golden.py runs the UNMODIFIED reference modules (through oracle/ref_shim.py, build
container only) on the same seeded inputs and asserts this restatement agrees; the resulting vectors are
committed under tests/golden/.  The reference has no sample-level golden vectors of its own for this path
(SURVEY.md section 4), so "parity pinned against the reference's own code run here", not against reference KATs.

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------------------------------------
# configuration


@dataclass
class MDXConfig:
    """Model / chunking constants (audio_separator/separator/architectures/mdx_separator.py:67-70,205-228)."""

    n_fft: int = 6144  # model_data["mdx_n_fft_scale_set"]
    hop_length: int = 1024  # arch_config["hop_length"]
    dim_f: int = 3072  # model_data["mdx_dim_f_set"]
    dim_t: int = 256  # 2 ** model_data["mdx_dim_t_set"]
    segment_size: int = 256  # arch_config["segment_size"]
    overlap: float = 0.25  # arch_config["overlap"]
    compensate: float = 1.022  # model_data["compensate"]
    enable_denoise: bool = False
    # ConvTDFNet hyper-parameters (uvr_lib_v5/mdxnet.py:30-38)
    dim_c: int = 4
    num_blocks: int = 11
    l: int = 3
    g: int = 48
    k: int = 3
    bn: int = 8

    @property
    def n_bins(self):  # mdx_separator.py:214
        return self.n_fft // 2 + 1

    @property
    def trim(self):  # mdx_separator.py:217
        return self.n_fft // 2

    @property
    def chunk_size(self):  # mdx_separator.py:220
        return self.hop_length * (self.segment_size - 1)

    @property
    def gen_size(self):  # mdx_separator.py:223
        return self.chunk_size - 2 * self.trim


# --------------------------------------------------------------------------------------------------------
# STFT / iSTFT  (audio_separator/separator/uvr_lib_v5/stft.py)


def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n, periodic=True) (stft.py:18), float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_forward(wave: np.ndarray, n_fft: int, hop: int, dim_f: int) -> np.ndarray:
    """STFT.__call__ (stft.py:20-56): (..., C, T) float32 -> (..., 2C, dim_f, T//hop+1) float32.

    torch.stft(center=True) reflect-pads n_fft//2 each side, frames at multiples of hop, multiplies by the
    periodic Hann window, one-sided un-normalised DFT.  Channel c maps to planes 2c (real) and 2c+1 (imag)
    (stft.py:44-50); the frequency axis is cropped to dim_f (stft.py:56).
    """
    wave = np.asarray(wave)
    lead = wave.shape[:-2]
    C, T = wave.shape[-2:]
    x = wave.reshape(-1, T).astype(np.float64)
    p = n_fft // 2
    xp = np.pad(x, ((0, 0), (p, p)), mode="reflect")
    n_frames = T // hop + 1
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = xp[:, idx] * hann_periodic(n_fft)  # (BC, frames, n_fft)
    spec = np.fft.rfft(frames, axis=-1)  # (BC, frames, n_fft/2+1)
    spec = np.transpose(spec, (0, 2, 1))[:, :dim_f, :]  # (BC, dim_f, frames)
    out = np.stack([spec.real, spec.imag], axis=1)  # (BC, 2, dim_f, frames)
    return out.reshape(*lead, C * 2, dim_f, n_frames).astype(np.float32)


def stft_inverse(spec: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """STFT.inverse (stft.py:99-126): (..., 2C, dim_f, frames) -> (..., C, hop*(frames-1)).

    Bins dim_f..n_fft/2 are zero-filled (stft.py:58-69); torch.istft(center=True) = irfft of each frame,
    times the window, overlap-added at hop, divided by the overlap-added squared window, with n_fft//2
    trimmed from both ends.
    """
    spec = np.asarray(spec)
    lead = spec.shape[:-3]
    C2, dim_f, n_frames = spec.shape[-3:]
    n_bins = n_fft // 2 + 1
    s = spec.reshape(-1, 2, dim_f, n_frames).astype(np.float64)
    z = np.zeros((s.shape[0], n_bins, n_frames), dtype=np.complex128)
    z[:, :dim_f, :] = s[:, 0] + 1j * s[:, 1]
    w = hann_periodic(n_fft)
    frames = np.fft.irfft(np.transpose(z, (0, 2, 1)), n=n_fft, axis=-1) * w  # (BC, frames, n_fft)
    full = n_fft + hop * (n_frames - 1)
    y = np.zeros((s.shape[0], full))
    env = np.zeros(full)
    for t in range(n_frames):
        y[:, t * hop : t * hop + n_fft] += frames[:, t]
        env[t * hop : t * hop + n_fft] += w * w
    p = n_fft // 2
    y = y[:, p : full - p] / env[p : full - p]
    return y.reshape(*lead, C2 // 2, hop * (n_frames - 1)).astype(np.float32)


# --------------------------------------------------------------------------------------------------------
# ConvTDFNet forward (the network inside UVR-MDX-NET-*.onnx; topology documented by
# audio_separator/separator/uvr_lib_v5/mdxnet.py:30-120 and modules.py:1-74)


def convtdfnet_param_shapes(cfg: MDXConfig):
    """Ordered (name, shape) list in the reference module's state_dict naming (mdxnet.py:53-98)."""
    g, l, k, bn = cfg.g, cfg.l, cfg.k, cfg.bn
    n = cfg.num_blocks // 2
    out = []

    def bn_(prefix, c):
        for nm in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{prefix}.{nm}", (c,)))

    def tfc_tdf(prefix, c, f):
        for i in range(l):
            out.append((f"{prefix}.tfc.H.{i}.0.weight", (c, c, k, k)))
            out.append((f"{prefix}.tfc.H.{i}.0.bias", (c,)))
            bn_(f"{prefix}.tfc.H.{i}.1", c)
        out.append((f"{prefix}.tdf.0.weight", (f // bn, f)))
        bn_(f"{prefix}.tdf.1", c)
        out.append((f"{prefix}.tdf.3.weight", (f, f // bn)))
        bn_(f"{prefix}.tdf.4", c)

    out.append(("first_conv.0.weight", (g, cfg.dim_c, 1, 1)))
    out.append(("first_conv.0.bias", (g,)))
    bn_("first_conv.1", g)
    f, c = cfg.dim_f, g
    for i in range(n):
        tfc_tdf(f"encoding_blocks.{i}", c, f)
        out.append((f"ds.{i}.0.weight", (c + g, c, 2, 2)))
        out.append((f"ds.{i}.0.bias", (c + g,)))
        bn_(f"ds.{i}.1", c + g)
        f //= 2
        c += g
    tfc_tdf("bottleneck_block", c, f)
    for i in range(n):
        out.append((f"us.{i}.0.weight", (c, c - g, 2, 2)))
        out.append((f"us.{i}.0.bias", (c - g,)))
        bn_(f"us.{i}.1", c - g)
        f *= 2
        c -= g
        tfc_tdf(f"decoding_blocks.{i}", c, f)
    out.append(("final_conv.0.weight", (cfg.dim_c, c, 1, 1)))
    out.append(("final_conv.0.bias", (cfg.dim_c,)))
    return out


def make_convtdfnet_weights(cfg: MDXConfig, seed: int = 0, out_gain: float = 1.0) -> dict:
    """Seeded synthetic weights (there are no real checkpoints offline).  He-style scaling keeps
    activations O(1) through the 11 blocks; BatchNorm statistics are non-trivial so that folding is tested."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in convtdfnet_param_shapes(cfg):
        if name.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("running_mean"):
            a = rng.normal(0.0, 0.1, shape)
        elif len(shape) == 1 and name.endswith(".weight"):  # BN gamma
            a = rng.uniform(0.8, 1.2, shape)
        elif len(shape) == 1:  # conv / BN bias
            a = rng.normal(0.0, 0.05, shape)
        elif len(shape) == 2:  # TDF linear (out, in)
            a = rng.normal(0.0, math.sqrt(2.0 / shape[1]), shape)
        elif name.startswith("us."):  # ConvTranspose2d weight (in, out, 2, 2): one tap contributes per output
            a = rng.normal(0.0, math.sqrt(2.0 / shape[0]), shape)
        else:  # Conv2d (out, in, kh, kw)
            a = rng.normal(0.0, math.sqrt(2.0 / (shape[1] * shape[2] * shape[3])), shape)
        # taming factors: the decoder multiplies by the skip (mdxnet.py:112), which squares magnitudes five times
        if name.startswith("us.") and name.endswith(".1.weight"):
            a = a * 0.2
        elif name.endswith(".tdf.4.weight"):
            a = a * 0.3
        elif name == "first_conv.0.weight":
            a = a * 0.05
        w[name] = a.astype(np.float32)
    w["final_conv.0.weight"] = (w["final_conv.0.weight"] * out_gain).astype(np.float32)
    return w


def convtdfnet_forward(weights: dict, cfg: MDXConfig, x: np.ndarray, dtype="float32") -> np.ndarray:
    """ConvTDFNet.forward (mdxnet.py:99-120) with TFC_TDF (modules.py:44-74), eval-mode BatchNorm.

    x: (B, 4, dim_f, dim_t) -> (B, 4, dim_f, dim_t).  torch-CPU functional ops, fp32 by default
    (dtype="float64" gives the rounding-free reference used to size tolerances).
    """
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k_: torch.from_numpy(np.asarray(v)).to(td) for k_, v in weights.items()}
    x = torch.from_numpy(np.ascontiguousarray(x)).to(td)
    n = cfg.num_blocks // 2

    def bn_relu(x, p):
        x = F.batch_norm(x, W[p + ".running_mean"], W[p + ".running_var"], W[p + ".weight"], W[p + ".bias"], False, 0.0, 1e-5)
        return F.relu(x)

    def tfc_tdf(x, p):
        for i in range(cfg.l):  # modules.py:20-23
            x = F.conv2d(x, W[f"{p}.tfc.H.{i}.0.weight"], W[f"{p}.tfc.H.{i}.0.bias"], padding=cfg.k // 2)
            x = bn_relu(x, f"{p}.tfc.H.{i}.1")
        t = F.linear(x, W[f"{p}.tdf.0.weight"])  # modules.py:63-70 (bias=False for the MDX nets)
        t = bn_relu(t, f"{p}.tdf.1")
        t = F.linear(t, W[f"{p}.tdf.3.weight"])
        t = bn_relu(t, f"{p}.tdf.4")
        return x + t  # modules.py:74

    with torch.no_grad():
        x = bn_relu(F.conv2d(x, W["first_conv.0.weight"], W["first_conv.0.bias"]), "first_conv.1")
        x = x.transpose(-1, -2)  # mdxnet.py:101
        skips = []
        for i in range(n):
            x = tfc_tdf(x, f"encoding_blocks.{i}")
            skips.append(x)
            x = bn_relu(F.conv2d(x, W[f"ds.{i}.0.weight"], W[f"ds.{i}.0.bias"], stride=2), f"ds.{i}.1")
        x = tfc_tdf(x, "bottleneck_block")
        for i in range(n):
            x = bn_relu(F.conv_transpose2d(x, W[f"us.{i}.0.weight"], W[f"us.{i}.0.bias"], stride=2), f"us.{i}.1")
            x = x * skips[-i - 1]  # mdxnet.py:112
            x = tfc_tdf(x, f"decoding_blocks.{i}")
        x = x.transpose(-1, -2)  # mdxnet.py:116
        x = F.conv2d(x, W["final_conv.0.weight"], W["final_conv.0.bias"])
    return x.to(torch.float32).numpy()


# --------------------------------------------------------------------------------------------------------
# run_model / demix / separate glue  (architectures/mdx_separator.py)


def run_model(mix_chunk: np.ndarray, cfg: MDXConfig, model_run, is_match_mix=False) -> np.ndarray:
    """MDXSeparator.run_model (mdx_separator.py:414-450): (B,2,chunk) -> (B,2,chunk)."""
    spek = stft_forward(mix_chunk, cfg.n_fft, cfg.hop_length, cfg.dim_f)
    spek[:, :, :3, :] *= 0  # mdx_separator.py:425
    if is_match_mix:
        spec_pred = spek  # mdx_separator.py:429-432
    elif cfg.enable_denoise:  # mdx_separator.py:435-440
        spec_pred = model_run(-spek) * -0.5 + model_run(spek) * 0.5
    else:
        spec_pred = model_run(spek)  # mdx_separator.py:443
    return stft_inverse(spec_pred, cfg.n_fft, cfg.hop_length)


def chunk_starts(n_samples: int, cfg: MDXConfig, is_match_mix=False):
    """Chunk grid of MDXSeparator.demix (mdx_separator.py:307-348): returns (L, step, starts)."""
    chunk = cfg.chunk_size
    overlap = 0.02 if is_match_mix else cfg.overlap  # mdx_separator.py:310-319
    gen = chunk - 2 * cfg.trim
    pad = gen + cfg.trim - (n_samples % gen)  # mdx_separator.py:327
    L = cfg.trim + n_samples + pad
    step = int((1 - overlap) * chunk)  # mdx_separator.py:335
    return L, step, list(range(0, L, step))


def demix(mix: np.ndarray, cfg: MDXConfig, model_run, is_match_mix=False, only_chunks=None) -> np.ndarray:
    """MDXSeparator.demix (mdx_separator.py:293-412): mix (2,N) float32 -> (2,N) float32.
    only_chunks (bench.py's bounded CPU sample): process just these indices of the chunk grid (the result is then only meaningful where they cover)."""
    mix = np.asarray(mix, dtype=np.float32)
    N = mix.shape[-1]
    chunk = cfg.chunk_size
    overlap = 0.02 if is_match_mix else cfg.overlap
    L, step, starts = chunk_starts(N, cfg, is_match_mix)
    mixture = np.zeros((2, L), dtype=np.float32)
    mixture[:, cfg.trim : cfg.trim + N] = mix  # mdx_separator.py:329
    result = np.zeros((1, 2, L), dtype=np.float32)
    divider = np.zeros((1, 2, L), dtype=np.float32)
    for ci, start in enumerate(starts):
        if only_chunks is not None and ci not in only_chunks:
            continue
        end = min(start + chunk, L)
        actual = end - start
        part = np.zeros((1, 2, chunk), dtype=np.float32)  # right zero-pad short last chunk, :363-366
        part[0, :, :actual] = mixture[:, start:end]
        tar = run_model(part, cfg, model_run, is_match_mix)
        if overlap != 0:
            window = np.hanning(actual)[None, None, :]  # symmetric Hann of the ACTUAL length, :358
            tar[..., :actual] = tar[..., :actual] * window
            divider[..., start:end] += window
        else:
            divider[..., start:end] += 1
        result[..., start:end] += tar[..., :actual]
    with np.errstate(divide="ignore", invalid="ignore"):
        tar_waves = result / divider  # mdx_separator.py:396 (0/0 -> nan at the window's end zeros, as in the reference)
    tar_waves = tar_waves[:, :, cfg.trim : -cfg.trim]  # :400
    return np.concatenate(tar_waves, axis=-1)[:, :N]  # :401


def normalize(wave: np.ndarray, max_peak=1.0, min_peak=None) -> np.ndarray:
    """spec_utils.normalize (uvr_lib_v5/spec_utils.py:99-115); returns a new array."""
    wave = np.array(wave, copy=True)
    maxv = np.abs(wave).max()
    if maxv > max_peak:
        wave *= max_peak / maxv
    elif min_peak is not None and maxv < min_peak:
        wave *= min_peak / maxv
    return wave


def separate_arrays(mix: np.ndarray, cfg: MDXConfig, model_run, normalization_threshold=0.9, amplification_threshold=0.0):
    """Array-level body of MDXSeparator.separate (mdx_separator.py:152-182), invert_using_spec=False.

    mix (2,N) float32 as returned by prepare_mix.  Returns (primary (N,2), secondary (N,2)) float32 --
    the arrays handed to write_audio, before its own per-stem normalise + int16 conversion.
    """
    mix = np.asarray(mix, dtype=np.float32)
    peak = np.abs(mix).max()  # :155
    mixn = normalize(mix, normalization_threshold, amplification_threshold)  # :156
    source = demix(mixn, cfg, model_run) * peak  # :159
    primary = source.T
    secondary = (-primary * cfg.compensate) + mixn.T  # :182
    return primary.astype(np.float32), secondary.astype(np.float32)


def to_pcm16(stem: np.ndarray, normalization_threshold=0.9, amplification_threshold=0.0):
    """write_audio_pydub's sample conversion (common_separator.py:310-339): normalise, near-silence
    check (returns None), (x*32767).astype(int16) truncation, L/R interleave."""
    s = normalize(stem, normalization_threshold, amplification_threshold)
    if np.max(np.abs(s)) < 1e-6:
        return None
    s16 = (s * 32767).astype(np.int16)
    out = np.empty((2 * s16.shape[0],), dtype=np.int16)
    out[0::2] = s16[:, 0]
    out[1::2] = s16[:, 1]
    return out


# --------------------------------------------------------------------------------------------------------
# synthetic programme material


def synth_music(n_samples: int, seed: int = 1234, sr: int = 44100) -> np.ndarray:
    """Deterministic band-limited 'music-like' stereo, float32 (2, n_samples), peak 0.95."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / sr
    out = np.zeros((2, n_samples))
    for ch in range(2):
        for _ in range(16):
            f0 = rng.uniform(55.0, 880.0)
            am = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.1, 2.0) * t + rng.uniform(0, 2 * np.pi))
            for kk in range(1, 6):
                if f0 * kk < sr / 2.2:
                    out[ch] += (am / kk) * np.sin(2 * np.pi * f0 * kk * t + rng.uniform(0, 2 * np.pi))
    white = rng.standard_normal((2, n_samples))
    spec = np.fft.rfft(white, axis=-1)
    fr = np.arange(spec.shape[-1], dtype=np.float64)
    fr[0] = 1.0
    pink = np.fft.irfft(spec / np.sqrt(fr), n=n_samples, axis=-1)
    pink /= np.abs(pink).max()
    gate = (np.sin(2 * np.pi * 2.0 * t) > 0.7).astype(np.float64)
    out = out / np.abs(out).max() + 0.1 * pink + 0.2 * gate * white / np.abs(white).max()
    mid = out.mean(0, keepdims=True)
    out = 0.7 * out + 0.3 * mid
    out *= 0.95 / np.abs(out).max()
    return out.astype(np.float32)
