"""CPU oracle for the MDXC (MDX23C, TFC_TDF_net) hot path -- TEST INFRASTRUCTURE, not product code.

Restates, with torch-CPU functional ops / numpy:
  * TFC_TDF_net.forward            audio_separator/separator/uvr_lib_v5/tfc_tdf_v3.py:230-267 (+ TFC_TDF :110-148,
                                   Upscale/Downscale :84-107, cac2cws/cws2cac :216-228, STFT :5-53)
  * MDXCSeparator.demix (non-Roformer branch)   audio_separator/separator/architectures/mdxc_separator.py:345-404
Pinned against the unmodified reference modules by oracle/make_golden_mdxc.py (build container only).
Only InstanceNorm(affine) + GELU networks (the MDX23C-8KFFT-InstVoc_HQ configuration) are covered.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

import mdx_oracle as M


@dataclass
class MDXCConfig:
    """model_2_stem_full_band_8k.yaml values by default (footnote)."""

    n_fft: int = 8192
    hop_length: int = 1024
    dim_f: int = 4096
    dim_t: int = 256  # inference.dim_t
    num_channels: int = 2
    num_subbands: int = 4
    num_scales: int = 5
    scale: tuple = (2, 2)
    num_blocks_per_scale: int = 2
    num_channels_model: int = 128  # model.num_channels
    growth: int = 128
    bottleneck_factor: int = 4
    instruments: tuple = ("Vocals", "Instrumental")
    target_instrument: str | None = None
    overlap: int = 8  # mdxc_params["overlap"]

    @property
    def num_targets(self):
        return 1 if self.target_instrument else len(self.instruments)

    @property
    def dim_c(self):
        return self.num_subbands * self.num_channels * 2

    @property
    def chunk_size(self):  # mdxc_separator.py:361
        return self.hop_length * (self.dim_t - 1)

    @property
    def hop_size(self):  # :364
        return self.chunk_size // self.overlap


def param_shapes(cfg: MDXCConfig):
    """(name, shape) in the reference module's state_dict order (tfc_tdf_v3.py:110-214)."""
    out = []
    n, l, c, g, bn = cfg.num_scales, cfg.num_blocks_per_scale, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor
    f = cfg.dim_f // cfg.num_subbands

    def norm(prefix, ch):
        out.append((f"{prefix}.weight", (ch,)))
        out.append((f"{prefix}.bias", (ch,)))

    def tfc_tdf(prefix, in_c, ch, ff):
        for i in range(l):
            b = f"{prefix}.blocks.{i}"
            norm(f"{b}.tfc1.0", in_c)
            out.append((f"{b}.tfc1.2.weight", (ch, in_c, 3, 3)))
            norm(f"{b}.tdf.0", ch)
            out.append((f"{b}.tdf.2.weight", (ff // bn, ff)))
            norm(f"{b}.tdf.3", ch)
            out.append((f"{b}.tdf.5.weight", (ff, ff // bn)))
            norm(f"{b}.tfc2.0", ch)
            out.append((f"{b}.tfc2.2.weight", (ch, ch, 3, 3)))
            out.append((f"{b}.shortcut.weight", (ch, in_c, 1, 1)))
            in_c = ch

    out.append(("first_conv.weight", (c, cfg.dim_c, 1, 1)))
    for i in range(n):
        tfc_tdf(f"encoder_blocks.{i}.tfc_tdf", c, c, f)
        norm(f"encoder_blocks.{i}.downscale.conv.0", c)
        out.append((f"encoder_blocks.{i}.downscale.conv.2.weight", (c + g, c, cfg.scale[0], cfg.scale[1])))
        f //= cfg.scale[1]
        c += g
    tfc_tdf("bottleneck_block", c, c, f)
    for i in range(n):
        norm(f"decoder_blocks.{i}.upscale.conv.0", c)
        out.append((f"decoder_blocks.{i}.upscale.conv.2.weight", (c, c - g, cfg.scale[0], cfg.scale[1])))
        f *= cfg.scale[1]
        c -= g
        tfc_tdf(f"decoder_blocks.{i}.tfc_tdf", 2 * c, c, f)
    out.append(("final_conv.0.weight", (c, c + cfg.dim_c, 1, 1)))
    out.append(("final_conv.2.weight", (cfg.num_targets * cfg.dim_c, c, 1, 1)))
    return out


def make_weights(cfg: MDXCConfig, seed=0, out_gain=1.0):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in param_shapes(cfg):
        if len(shape) == 1 and name.endswith(".weight"):
            a = rng.uniform(0.7, 1.3, shape)  # InstanceNorm gamma
        elif len(shape) == 1:
            a = rng.normal(0.0, 0.1, shape)  # InstanceNorm beta
        elif len(shape) == 2:
            a = rng.normal(0.0, math.sqrt(1.0 / shape[1]), shape)
        elif "upscale" in name:
            a = rng.normal(0.0, math.sqrt(1.0 / shape[0]), shape)
        else:
            a = rng.normal(0.0, math.sqrt(1.0 / (shape[1] * shape[2] * shape[3])), shape)
        w[name] = a.astype(np.float32)
    w["first_conv.weight"] = (w["first_conv.weight"] * 0.1).astype(np.float32)
    w["final_conv.2.weight"] = (w["final_conv.2.weight"] * out_gain).astype(np.float32)
    return w


def net_forward_spec(weights, cfg: MDXCConfig, spec: np.ndarray, dtype="float32") -> np.ndarray:
    """The part of TFC_TDF_net.forward between the two STFTs (tfc_tdf_v3.py:234-261):
    spec (B, 4, dim_f, T) -> (B, S*4, dim_f, T)  [S = num_targets; the reshape to (B,S,4,...) is the caller's]."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items()}
    x = torch.from_numpy(np.ascontiguousarray(spec)).to(td)
    k = cfg.num_subbands
    l = cfg.num_blocks_per_scale

    def na(x, p):  # norm(c) + act  (InstanceNorm2d(affine) + GELU)
        return F.gelu(F.instance_norm(x, None, None, W[p + ".weight"], W[p + ".bias"], True, 0.0, 1e-5))

    def tfc_tdf(x, prefix):
        for i in range(l):
            b = f"{prefix}.blocks.{i}"
            s = F.conv2d(x, W[f"{b}.shortcut.weight"])
            x = F.conv2d(na(x, f"{b}.tfc1.0"), W[f"{b}.tfc1.2.weight"], padding=1)
            t = F.linear(na(x, f"{b}.tdf.0"), W[f"{b}.tdf.2.weight"])
            t = F.linear(na(t, f"{b}.tdf.3"), W[f"{b}.tdf.5.weight"])
            x = x + t
            x = F.conv2d(na(x, f"{b}.tfc2.0"), W[f"{b}.tfc2.2.weight"], padding=1)
            x = x + s
        return x

    with torch.no_grad():
        b_, c_, f_, t_ = x.shape
        x = x.reshape(b_, c_ * k, f_ // k, t_)  # cac2cws :216-221
        mix = x
        first = x = F.conv2d(x, W["first_conv.weight"])
        x = x.transpose(-1, -2)
        enc = []
        for i in range(cfg.num_scales):
            x = tfc_tdf(x, f"encoder_blocks.{i}.tfc_tdf")
            enc.append(x)
            x = F.conv2d(na(x, f"encoder_blocks.{i}.downscale.conv.0"), W[f"encoder_blocks.{i}.downscale.conv.2.weight"], stride=cfg.scale)
        x = tfc_tdf(x, "bottleneck_block")
        for i in range(cfg.num_scales):
            x = F.conv_transpose2d(na(x, f"decoder_blocks.{i}.upscale.conv.0"), W[f"decoder_blocks.{i}.upscale.conv.2.weight"], stride=cfg.scale)
            x = torch.cat([x, enc.pop()], 1)
            x = tfc_tdf(x, f"decoder_blocks.{i}.tfc_tdf")
        x = x.transpose(-1, -2)
        x = x * first  # :255
        x = F.conv2d(torch.cat([mix, x], 1), W["final_conv.0.weight"])
        x = F.conv2d(F.gelu(x), W["final_conv.2.weight"])
        b_, c_, f_, t_ = x.shape
        x = x.reshape(b_, c_ // k, f_ * k, t_)  # cws2cac :223-228
    return x.to(torch.float32).numpy()


def net_forward(weights, cfg: MDXCConfig, wave: np.ndarray, dtype="float32") -> np.ndarray:
    """TFC_TDF_net.forward: (B, 2, chunk) -> (B, S, 2, chunk) (S squeezed away by the reference when S == 1)."""
    spec = M.stft_forward(wave, cfg.n_fft, cfg.hop_length, cfg.dim_f)
    y = net_forward_spec(weights, cfg, spec, dtype)
    B = y.shape[0]
    S = cfg.num_targets
    y = y.reshape(B, S, 4, cfg.dim_f, y.shape[-1])
    out = M.stft_inverse(y, cfg.n_fft, cfg.hop_length)  # (B, S, 2, chunk)
    return out[:, 0] if S == 1 else out


def demix(mix: np.ndarray, cfg: MDXCConfig, model_run) -> np.ndarray:
    """MDXCSeparator.demix, non-Roformer branch (mdxc_separator.py:345-404): (2, N) -> (S, 2, N) (or (2, N) for S == 1).
    model_run: (B, 2, chunk) -> (B, S, 2, chunk)."""
    mix = np.asarray(mix, dtype=np.float32)
    N = mix.shape[1]
    chunk, hop = cfg.chunk_size, cfg.hop_size
    pad = hop - (N - chunk) % hop  # :367
    front = chunk - hop
    padded = np.concatenate([np.zeros((2, front), np.float32), mix, np.zeros((2, pad + front), np.float32)], 1)  # :371
    Lp = padded.shape[1]
    n_chunks = (Lp - chunk) // hop + 1  # unfold(1, chunk, hop)
    S = cfg.num_targets
    acc = np.zeros((S, 2, Lp), np.float32) if S > 1 else np.zeros((2, Lp), np.float32)
    for i in range(n_chunks):
        out = model_run(padded[None, :, i * hop : i * hop + chunk])[0]
        acc[..., i * hop : i * hop + chunk] += out
    return acc[..., front : Lp - (pad + front)] / cfg.overlap  # :402


def chunk_grid(n_samples: int, cfg: MDXCConfig):
    chunk, hop = cfg.chunk_size, cfg.hop_size
    pad = hop - (n_samples - chunk) % hop
    front = chunk - hop
    Lp = front + n_samples + pad + front
    return Lp, front, pad, (Lp - chunk) // hop + 1
