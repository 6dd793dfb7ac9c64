"""CPU oracle for the BS-Roformer hot path (the Roformer branch of the MDXC plugin) -- TEST INFRASTRUCTURE, not product code.

Restates, with torch-CPU functional ops / numpy:
  * BSRoformer.forward            uvr_lib_v5/roformer/bs_roformer.py:418-497 (+ RMSNorm :30-37, FeedForward :40-48, Attention :51-82,
                                  Transformer :112-131, BandSplit :134-150, MLP/MaskEstimator :153-190, DEFAULT_FREQS_PER_BANDS :193-256)
  * Attend.forward                uvr_lib_v5/roformer/attend.py:85-112 (softmax(q k^T d^-1/2) v; the flash path is the same function)
  * MDXCSeparator.demix, Roformer branch   architectures/mdxc_separator.py:272-343 (+ overlap_add :246-255, stem dict :406-468)
Third-party piece restated from its published definition: rotary-embedding-torch (lucidrains, the `RotaryEmbedding(dim=dim_head)` defaults:
freqs_for="lang", theta=10000, interleaved (d r) pairing, rotate_half = (-x2, x1)) -- it is NOT installed here, so the reference is pinned
with THIS restatement in its place (oracle/make_golden_roformer.py); everything else is the unmodified reference.
Covered structure: linear_transformer_depth = 0, no sage attention, eval mode (dropout off).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

DEFAULT_FREQS_PER_BANDS = (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)


@dataclass
class BSRoformerConfig:
    """model section of a BS-Roformer YAML; defaults = model_bs_roformer_ep_317_sdr_12.9755 (the reference's default model)."""

    dim: int = 512
    depth: int = 12
    stereo: bool = True
    num_stems: int = 1
    time_transformer_depth: int = 1
    freq_transformer_depth: int = 1
    freqs_per_bands: tuple = DEFAULT_FREQS_PER_BANDS
    dim_head: int = 64
    heads: int = 8
    mlp_expansion_factor: int = 4
    mask_estimator_depth: int = 2
    stft_n_fft: int = 2048
    stft_hop_length: int = 512  # NOTE the released configs use 441; 512 keeps synthetic tests on a power-of-two grid
    stft_win_length: int = 2048
    dim_t: int = 801  # inference.dim_t
    sample_rate: int = 44100
    overlap: int = 8  # mdxc_params["overlap"] (seconds for the Roformer branch)

    @property
    def audio_channels(self):
        return 2 if self.stereo else 1

    @property
    def chunk_size(self):  # mdxc_separator.py:301
        return self.stft_hop_length * (self.dim_t - 1)

    @property
    def step(self):  # :308-309
        desired = int(self.overlap * self.sample_rate)
        return self.chunk_size if desired <= 0 else min(desired, self.chunk_size)

    @property
    def band_dims(self):
        return tuple(2 * f * self.audio_channels for f in self.freqs_per_bands)

    def kwargs(self):
        return dict(dim=self.dim, depth=self.depth, stereo=self.stereo, num_stems=self.num_stems, time_transformer_depth=self.time_transformer_depth,
                    freq_transformer_depth=self.freq_transformer_depth, freqs_per_bands=tuple(self.freqs_per_bands), dim_head=self.dim_head, heads=self.heads,
                    mlp_expansion_factor=self.mlp_expansion_factor, mask_estimator_depth=self.mask_estimator_depth, stft_n_fft=self.stft_n_fft,
                    stft_hop_length=self.stft_hop_length, stft_win_length=self.stft_win_length, flash_attn=False)


def param_shapes(cfg: BSRoformerConfig):
    out = []
    inner = cfg.heads * cfg.dim_head
    ff = int(cfg.dim * 4)
    for i in range(cfg.depth):
        for j, tdepth in enumerate((cfg.time_transformer_depth, cfg.freq_transformer_depth)):
            for l in range(tdepth):
                p = f"layers.{i}.{j}.layers.{l}"
                out.extend([(f"{p}.0.rotary_embed.freqs", (cfg.dim_head // 2,)), (f"{p}.0.norm.gamma", (cfg.dim,)), (f"{p}.0.to_qkv.weight", (3 * inner, cfg.dim)),
                            (f"{p}.0.to_gates.weight", (cfg.heads, cfg.dim)), (f"{p}.0.to_gates.bias", (cfg.heads,)), (f"{p}.0.to_out.0.weight", (cfg.dim, inner)),
                            (f"{p}.1.net.0.gamma", (cfg.dim,)), (f"{p}.1.net.1.weight", (ff, cfg.dim)), (f"{p}.1.net.1.bias", (ff,)), (f"{p}.1.net.4.weight", (cfg.dim, ff)),
                            (f"{p}.1.net.4.bias", (cfg.dim,))])
    out.append(("final_norm.gamma", (cfg.dim,)))
    for b, d_in in enumerate(cfg.band_dims):
        out.extend([(f"band_split.to_features.{b}.0.gamma", (d_in,)), (f"band_split.to_features.{b}.1.weight", (cfg.dim, d_in)), (f"band_split.to_features.{b}.1.bias", (cfg.dim,))])
    hid = cfg.dim * cfg.mlp_expansion_factor
    for s in range(cfg.num_stems):
        for b, d_in in enumerate(cfg.band_dims):
            dims = (cfg.dim,) + (hid,) * (cfg.mask_estimator_depth - 1) + (2 * d_in,)
            for li, (a, c) in enumerate(zip(dims[:-1], dims[1:])):
                out.extend([(f"mask_estimators.{s}.to_freqs.{b}.0.{2 * li}.weight", (c, a)), (f"mask_estimators.{s}.to_freqs.{b}.0.{2 * li}.bias", (c,))])
    return out


def rotary_freqs(dim_head: int, theta=10000.0) -> np.ndarray:
    """RotaryEmbedding(dim).freqs for freqs_for="lang": 1 / theta^(arange(0, dim, 2) / dim)."""
    return (1.0 / (theta ** (np.arange(0, dim_head, 2, dtype=np.float32)[: dim_head // 2] / dim_head))).astype(np.float32)


def make_weights(cfg: BSRoformerConfig, seed=0):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in param_shapes(cfg):
        if name.endswith("rotary_embed.freqs"):
            a = rotary_freqs(cfg.dim_head)
        elif name.endswith("gamma"):
            a = rng.uniform(0.7, 1.3, shape)
        elif name.endswith("bias"):
            a = rng.normal(0.0, 0.05, shape)
        else:
            a = rng.normal(0.0, math.sqrt(1.0 / shape[1]), shape)
        w[name] = np.asarray(a, dtype=np.float32)
    return w


def apply_rotary(t, freqs):
    """rotate_queries_or_keys(t) for t (..., n, d): positions arange(n), interleaved pairs (x[2i], x[2i+1]) rotated by pos * freqs[i]."""
    import torch

    n, d = t.shape[-2], t.shape[-1]
    ang = torch.arange(n, dtype=freqs.dtype)[:, None] * freqs[None, :]  # (n, d/2)
    ang = ang.repeat_interleave(2, dim=-1)  # "... n -> ... (n r)", r = 2
    x = t.reshape(*t.shape[:-1], d // 2, 2)
    rot = torch.stack((-x[..., 1], x[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def forward(weights, cfg: BSRoformerConfig, raw_audio: np.ndarray, dtype="float32") -> np.ndarray:
    """BSRoformer.forward (eval, target=None): (b, s, t) -> (b, s, t') for num_stems == 1, else (b, n, s, t')."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items()}
    x_in = torch.from_numpy(np.ascontiguousarray(raw_audio)).to(td)
    b, s, _ = x_in.shape
    H, dh = cfg.heads, cfg.dim_head

    def rms(x, g):
        return F.normalize(x, dim=-1) * (x.shape[-1] ** 0.5) * g

    def attention(x, p):
        xn = rms(x, W[f"{p}.norm.gamma"])
        qkv = F.linear(xn, W[f"{p}.to_qkv.weight"])
        B_, n, _ = qkv.shape
        q, k, v = qkv.view(B_, n, 3, H, dh).permute(2, 0, 3, 1, 4)
        fr = W[f"{p}.rotary_embed.freqs"]
        q, k = apply_rotary(q, fr), apply_rotary(k, fr)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh**-0.5
        out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
        gates = F.linear(xn, W[f"{p}.to_gates.weight"], W[f"{p}.to_gates.bias"])  # (B, n, H)
        out = out * gates.permute(0, 2, 1)[..., None].sigmoid()
        return F.linear(out.permute(0, 2, 1, 3).reshape(B_, n, H * dh), W[f"{p}.to_out.0.weight"])

    def feedforward(x, p):
        h = F.gelu(F.linear(rms(x, W[f"{p}.net.0.gamma"]), W[f"{p}.net.1.weight"], W[f"{p}.net.1.bias"]))
        return F.linear(h, W[f"{p}.net.4.weight"], W[f"{p}.net.4.bias"])

    def transformer(x, p, depth):
        for l in range(depth):
            x = attention(x, f"{p}.layers.{l}.0") + x
            x = feedforward(x, f"{p}.layers.{l}.1") + x
        return x

    with torch.no_grad():
        win = torch.hann_window(cfg.stft_win_length).to(td)
        st = torch.stft(x_in.reshape(b * s, -1), cfg.stft_n_fft, cfg.stft_hop_length, cfg.stft_win_length, window=win, normalized=False, return_complex=True)
        st = torch.view_as_real(st).view(b, s, st.shape[-2], st.shape[-1], 2)
        Fq, T = st.shape[2], st.shape[3]
        stft_repr = st.permute(0, 2, 1, 3, 4).reshape(b, Fq * s, T, 2)  # "b s f t c -> b (f s) t c"
        x = stft_repr.permute(0, 2, 1, 3).reshape(b, T, Fq * s * 2)  # "b f t c -> b t (f c)"
        feats, off = [], 0
        for bi, d_in in enumerate(cfg.band_dims):
            xb = x[..., off : off + d_in]
            off += d_in
            feats.append(F.linear(rms(xb, W[f"band_split.to_features.{bi}.0.gamma"]), W[f"band_split.to_features.{bi}.1.weight"], W[f"band_split.to_features.{bi}.1.bias"]))
        x = torch.stack(feats, dim=-2)  # (b, t, nb, d)
        nb = x.shape[2]
        for i in range(cfg.depth):
            x = x.permute(0, 2, 1, 3).reshape(b * nb, T, cfg.dim)
            x = transformer(x, f"layers.{i}.0", cfg.time_transformer_depth)
            x = x.view(b, nb, T, cfg.dim).permute(0, 2, 1, 3).reshape(b * T, nb, cfg.dim)
            x = transformer(x, f"layers.{i}.1", cfg.freq_transformer_depth)
            x = x.view(b, T, nb, cfg.dim)
        x = rms(x, W["final_norm.gamma"])
        masks = []
        for si in range(cfg.num_stems):
            outs = []
            for bi in range(nb):
                h = x[:, :, bi]
                nl = cfg.mask_estimator_depth
                for li in range(nl):
                    p = f"mask_estimators.{si}.to_freqs.{bi}.0.{2 * li}"
                    h = F.linear(h, W[f"{p}.weight"], W[f"{p}.bias"])
                    if li < nl - 1:
                        h = torch.tanh(h)
                outs.append(F.glu(h, dim=-1))
            masks.append(torch.cat(outs, dim=-1))
        mask = torch.stack(masks, dim=1)  # (b, n, t, F*s*2)
        mask = mask.view(b, cfg.num_stems, T, Fq * s, 2).permute(0, 1, 3, 2, 4)  # "b n t (f c) -> b n f t c"
        sc = torch.view_as_complex(stft_repr.contiguous())[:, None] * torch.view_as_complex(mask.contiguous())  # (b, n, F*s, T)
        sc = sc.view(b, cfg.num_stems, Fq, s, T).permute(0, 1, 3, 2, 4).reshape(b * cfg.num_stems * s, Fq, T)  # "b n (f s) t -> (b n s) f t"
        rec = torch.istft(sc, cfg.stft_n_fft, cfg.stft_hop_length, cfg.stft_win_length, window=win, normalized=False, return_complex=False)
        rec = rec.view(b, cfg.num_stems, s, -1)
        if cfg.num_stems == 1:
            rec = rec[:, 0]
    return rec.to(torch.float32).numpy()


def chunk_starts(n_samples: int, cfg: BSRoformerConfig):
    return list(range(0, n_samples, cfg.step))


def demix(mix: np.ndarray, cfg: BSRoformerConfig, model_run, n_instruments=None) -> np.ndarray:
    """MDXCSeparator.demix, Roformer branch (mdxc_separator.py:272-343): mix (2, N) -> (S, 2, N), S = len(training.instruments).
    model_run: (1, 2, L) -> (1, 2, L') or (1, n, 2, L')."""
    from scipy import signal

    mix = np.asarray(mix, dtype=np.float32)
    N = mix.shape[1]
    C, step = cfg.chunk_size, cfg.step
    S = n_instruments or max(1, cfg.num_stems)
    window = signal.windows.hamming(C).astype(np.float32)
    result = np.zeros((S, 2, N), np.float32)
    counter = np.zeros((S, 2, N), np.float32)
    for i in range(0, N, step):
        part = mix[:, i : i + C]
        length = part.shape[-1]
        tail = i + C > N
        if tail:
            part = mix[:, -C:]
            length = C
        x = model_run(part[None])[0]
        start = N - C if tail else i
        safe = min(length, x.shape[-1], C)
        if safe > 0:
            result[..., start : start + safe] += x[..., :safe] * window[:safe]
            counter[..., start : start + safe] += window[:safe]
    return result / np.maximum(counter, 1e-10)


# =========================================================================================================================
# Mel-Band Roformer (uvr_lib_v5/roformer/mel_band_roformer.py): same transformer stack, but (a) the bands are the supports of a mel
# filter bank and OVERLAP (features are gathered per band with `freq_indices`, masks are scatter-added and averaged by the number of bands
# covering each frequency, :239-262, :300-318), (b) every Transformer ends with an RMSNorm (norm_output=True) and there is no final_norm,
# (c) MLP(depth) has `depth` hidden layers (:106-117).  librosa.filters.mel (absent here) only contributes the SUPPORT of each filter;
# its Slaney mel scale is restated below from librosa's published source -- PARITY UNPINNED for the band layout itself (the reference is
# pinned with this restatement injected as librosa.filters.mel).
@dataclass
class MelBandRoformerConfig:
    dim: int = 384
    depth: int = 6
    stereo: bool = True
    num_stems: int = 1
    time_transformer_depth: int = 1
    freq_transformer_depth: int = 1
    num_bands: int = 60
    dim_head: int = 64
    heads: int = 8
    mask_estimator_depth: int = 2
    sample_rate: int = 44100
    stft_n_fft: int = 2048
    stft_hop_length: int = 441
    stft_win_length: int = 2048
    dim_t: int = 801
    overlap: int = 8

    @property
    def audio_channels(self):
        return 2 if self.stereo else 1

    @property
    def chunk_size(self):
        return self.stft_hop_length * (self.dim_t - 1)

    @property
    def step(self):
        desired = int(self.overlap * self.sample_rate)
        return self.chunk_size if desired <= 0 else min(desired, self.chunk_size)

    def kwargs(self):
        return dict(dim=self.dim, depth=self.depth, stereo=self.stereo, num_stems=self.num_stems, time_transformer_depth=self.time_transformer_depth,
                    freq_transformer_depth=self.freq_transformer_depth, num_bands=self.num_bands, dim_head=self.dim_head, heads=self.heads,
                    mask_estimator_depth=self.mask_estimator_depth, sample_rate=self.sample_rate, stft_n_fft=self.stft_n_fft, stft_hop_length=self.stft_hop_length,
                    stft_win_length=self.stft_win_length, flash_attn=False)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    if f.ndim:
        t = f >= min_log_hz
        mels[t] = min_log_mel + np.log(f[t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    t = m >= min_log_mel
    freqs[t] = min_log_hz * np.exp(logstep * (m[t] - min_log_mel))
    return freqs


def mel_filter_bank(sr, n_fft, n_mels):
    """librosa.filters.mel(sr=, n_fft=, n_mels=) with its defaults (fmin 0, fmax sr/2, Slaney scale, norm="slaney", float32)."""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def mel_band_layout(cfg: MelBandRoformerConfig):
    """-> (freqs_per_band bool (bands, F), freq_indices int64 over the (f s) axis, num_freqs_per_band, num_bands_per_freq)  (:239-262)"""
    fb = mel_filter_bank(cfg.sample_rate, cfg.stft_n_fft, cfg.num_bands).copy()
    fb[0][0] = 1.0
    fb[-1, -1] = 1.0
    fpb = fb > 0
    assert fpb.any(axis=0).all(), "all frequencies need to be covered by all bands for now"
    Fq = fpb.shape[1]
    idx = np.tile(np.arange(Fq), (cfg.num_bands, 1))[fpb]
    if cfg.stereo:
        idx = (idx[:, None] * 2 + np.arange(2)[None, :]).reshape(-1)
    return fpb, idx.astype(np.int64), fpb.sum(1), fpb.sum(0)


def mel_band_dims(cfg: MelBandRoformerConfig):
    return tuple(int(2 * f * cfg.audio_channels) for f in mel_band_layout(cfg)[2])


def mel_param_shapes(cfg: MelBandRoformerConfig):
    out = []
    inner = cfg.heads * cfg.dim_head
    ff = int(cfg.dim * 4)
    for i in range(cfg.depth):
        for j, tdepth in enumerate((cfg.time_transformer_depth, cfg.freq_transformer_depth)):
            for l in range(tdepth):
                p = f"layers.{i}.{j}.layers.{l}"
                out.extend([(f"{p}.0.rotary_embed.freqs", (cfg.dim_head // 2,)), (f"{p}.0.norm.gamma", (cfg.dim,)), (f"{p}.0.to_qkv.weight", (3 * inner, cfg.dim)),
                            (f"{p}.0.to_gates.weight", (cfg.heads, cfg.dim)), (f"{p}.0.to_gates.bias", (cfg.heads,)), (f"{p}.0.to_out.0.weight", (cfg.dim, inner)),
                            (f"{p}.1.net.0.gamma", (cfg.dim,)), (f"{p}.1.net.1.weight", (ff, cfg.dim)), (f"{p}.1.net.1.bias", (ff,)), (f"{p}.1.net.4.weight", (cfg.dim, ff)),
                            (f"{p}.1.net.4.bias", (cfg.dim,))])
            out.append((f"layers.{i}.{j}.norm.gamma", (cfg.dim,)))
    dims = mel_band_dims(cfg)
    for b, d_in in enumerate(dims):
        out.extend([(f"band_split.to_features.{b}.0.gamma", (d_in,)), (f"band_split.to_features.{b}.1.weight", (cfg.dim, d_in)), (f"band_split.to_features.{b}.1.bias", (cfg.dim,))])
    hid = cfg.dim * 4
    for s in range(cfg.num_stems):
        for b, d_in in enumerate(dims):
            chain = (cfg.dim,) + (hid,) * cfg.mask_estimator_depth + (2 * d_in,)
            for li, (a, c) in enumerate(zip(chain[:-1], chain[1:])):
                out.extend([(f"mask_estimators.{s}.to_freqs.{b}.0.{2 * li}.weight", (c, a)), (f"mask_estimators.{s}.to_freqs.{b}.0.{2 * li}.bias", (c,))])
    return out


def make_mel_weights(cfg: MelBandRoformerConfig, seed=0):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in mel_param_shapes(cfg):
        if name.endswith("rotary_embed.freqs"):
            a = rotary_freqs(cfg.dim_head)
        elif name.endswith("gamma"):
            a = rng.uniform(0.7, 1.3, shape)
        elif name.endswith("bias"):
            a = rng.normal(0.0, 0.05, shape)
        else:
            a = rng.normal(0.0, math.sqrt(1.0 / shape[1]), shape)
        w[name] = np.asarray(a, dtype=np.float32)
    return w


def forward_mel(weights, cfg: MelBandRoformerConfig, raw_audio: np.ndarray, dtype="float32") -> np.ndarray:
    """MelBandRoformer.forward (eval, target=None, match_input_audio_length=False): (b, s, t) -> (b, s, t') or (b, n, s, t')."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items()}
    x_in = torch.from_numpy(np.ascontiguousarray(raw_audio)).to(td)
    b, s, _ = x_in.shape
    H, dh = cfg.heads, cfg.dim_head
    fpb, freq_indices, nfpb, nbpf = mel_band_layout(cfg)
    fi = torch.from_numpy(freq_indices)
    dims = [int(2 * f * s) for f in nfpb]

    def rms(x, g):
        return F.normalize(x, dim=-1) * (x.shape[-1] ** 0.5) * g

    def attention(x, p):
        xn = rms(x, W[f"{p}.norm.gamma"])
        qkv = F.linear(xn, W[f"{p}.to_qkv.weight"])
        B_, n, _ = qkv.shape
        q, k, v = qkv.view(B_, n, 3, H, dh).permute(2, 0, 3, 1, 4)
        fr = W[f"{p}.rotary_embed.freqs"]
        q, k = apply_rotary(q, fr), apply_rotary(k, fr)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh**-0.5
        out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
        gates = F.linear(xn, W[f"{p}.to_gates.weight"], W[f"{p}.to_gates.bias"])
        out = out * gates.permute(0, 2, 1)[..., None].sigmoid()
        return F.linear(out.permute(0, 2, 1, 3).reshape(B_, n, H * dh), W[f"{p}.to_out.0.weight"])

    def transformer(x, p, depth):
        for l in range(depth):
            x = attention(x, f"{p}.layers.{l}.0") + x
            q_ = f"{p}.layers.{l}.1"
            h = F.gelu(F.linear(rms(x, W[f"{q_}.net.0.gamma"]), W[f"{q_}.net.1.weight"], W[f"{q_}.net.1.bias"]))
            x = F.linear(h, W[f"{q_}.net.4.weight"], W[f"{q_}.net.4.bias"]) + x
        return rms(x, W[f"{p}.norm.gamma"])

    with torch.no_grad():
        win = torch.hann_window(cfg.stft_win_length).to(td)
        st = torch.stft(x_in.reshape(b * s, -1), cfg.stft_n_fft, cfg.stft_hop_length, cfg.stft_win_length, window=win, normalized=False, return_complex=True)
        st = torch.view_as_real(st).view(b, s, st.shape[-2], st.shape[-1], 2)
        Fq, T = st.shape[2], st.shape[3]
        stft_repr = st.permute(0, 2, 1, 3, 4).reshape(b, Fq * s, T, 2)
        x = stft_repr[:, fi]  # (b, G, T, 2), G = sum over bands of their (freq, channel) pairs
        x = x.permute(0, 2, 1, 3).reshape(b, T, -1)
        feats, off = [], 0
        for bi, d_in in enumerate(dims):
            feats.append(F.linear(rms(x[..., off : off + d_in], W[f"band_split.to_features.{bi}.0.gamma"]), W[f"band_split.to_features.{bi}.1.weight"], W[f"band_split.to_features.{bi}.1.bias"]))
            off += d_in
        x = torch.stack(feats, dim=-2)
        nb = x.shape[2]
        for i in range(cfg.depth):
            x = x.permute(0, 2, 1, 3).reshape(b * nb, T, cfg.dim)
            x = transformer(x, f"layers.{i}.0", cfg.time_transformer_depth)
            x = x.view(b, nb, T, cfg.dim).permute(0, 2, 1, 3).reshape(b * T, nb, cfg.dim)
            x = transformer(x, f"layers.{i}.1", cfg.freq_transformer_depth)
            x = x.view(b, T, nb, cfg.dim)
        masks = []
        n_lin = cfg.mask_estimator_depth + 1
        for si in range(cfg.num_stems):
            outs = []
            for bi in range(nb):
                h = x[:, :, bi]
                for li in range(n_lin):
                    p = f"mask_estimators.{si}.to_freqs.{bi}.0.{2 * li}"
                    h = F.linear(h, W[f"{p}.weight"], W[f"{p}.bias"])
                    if li < n_lin - 1:
                        h = torch.tanh(h)
                outs.append(F.glu(h, dim=-1))
            masks.append(torch.cat(outs, dim=-1))
        masks = torch.stack(masks, dim=1)  # (b, n, t, G*2)
        G = fi.numel()
        masks = torch.view_as_complex(masks.view(b, cfg.num_stems, T, G, 2).permute(0, 1, 3, 2, 4).contiguous())  # (b, n, G, t)
        sc = torch.view_as_complex(stft_repr.contiguous())[:, None].expand(b, cfg.num_stems, Fq * s, T)
        summed = torch.zeros((b, cfg.num_stems, Fq * s, T), dtype=masks.dtype).scatter_add_(2, fi[None, None, :, None].expand(b, cfg.num_stems, G, T), masks)
        denom = torch.from_numpy(np.repeat(nbpf, s)).to(td)[:, None].clamp(min=1e-8)
        out = sc * (summed / denom)
        out = out.view(b, cfg.num_stems, Fq, s, T).permute(0, 1, 3, 2, 4).reshape(b * cfg.num_stems * s, Fq, T)
        rec = torch.istft(out, cfg.stft_n_fft, cfg.stft_hop_length, cfg.stft_win_length, window=win, normalized=False, return_complex=False)
        rec = rec.view(b, cfg.num_stems, s, -1)
        if cfg.num_stems == 1:
            rec = rec[:, 0]
    return rec.to(torch.float32).numpy()
