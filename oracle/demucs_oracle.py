"""CPU oracle for the Demucs (HTDemucs v4) hot path -- TEST INFRASTRUCTURE, not product code.

Functional torch-CPU / numpy restatement of
  * HTDemucs.forward                 uvr_lib_v5/demucs/htdemucs.py:483-620 (+ _spec :383-403, _ispec :405-413, _magnitude, _mask)
  * HEncLayer / HDecLayer            uvr_lib_v5/demucs/hdemucs.py:67-153, :252-330
  * DConv / LayerScale               uvr_lib_v5/demucs/demucs.py:85-168
  * CrossTransformerEncoder + layers uvr_lib_v5/demucs/transformer.py:19-49, :196-409, :529-560
  * apply_model / TensorChunk / center_trim   uvr_lib_v5/demucs/apply.py:71-113,124-260, utils.py:53-70
  * DemucsSeparator.demix_demucs     architectures/demucs_separator.py:162-195
for the structure of the released htdemucs models: depth 4, nfft 4096, no GroupNorm inside the encoder/decoder layers
(norm_starts >= depth), complex-as-channels, DConv in encoder and decoder (dconv_mode 3), sin embeddings, norm_first + layer
scale + norm_out transformer, cross_first False.  Pinned against the unmodified reference by oracle/make_golden_demucs.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction

import numpy as np


@dataclass
class HTConfig:
    sources: tuple = ("drums", "bass", "other", "vocals")
    audio_channels: int = 2
    channels: int = 48
    growth: int = 2
    nfft: int = 4096
    depth: int = 4
    kernel_size: int = 8
    stride: int = 4
    context: int = 1
    dconv_mode: int = 3  # bit 0: DConv in the encoders, bit 1: in the decoders (hdemucs.py:83-84, :268-269)
    dconv_depth: int = 2
    dconv_comp: int = 8
    bottom_channels: int = 512
    t_layers: int = 5
    t_heads: int = 8
    t_hidden_scale: float = 4.0
    freq_emb: float = 0.2
    emb_scale: float = 10.0
    samplerate: int = 44100
    segment: Fraction = Fraction(39, 5)
    max_period: float = 10000.0

    @property
    def hop(self):
        return self.nfft // 4

    @property
    def seg_len(self):
        return int(self.segment * self.samplerate)

    def kwargs(self):  # what the reference constructor receives
        return dict(sources=list(self.sources), audio_channels=self.audio_channels, channels=self.channels, growth=self.growth, nfft=self.nfft, depth=self.depth,
                    kernel_size=self.kernel_size, stride=self.stride, context=self.context, dconv_mode=self.dconv_mode, dconv_depth=self.dconv_depth, dconv_comp=self.dconv_comp,
                    bottom_channels=self.bottom_channels, t_layers=self.t_layers, t_heads=self.t_heads, t_hidden_scale=self.t_hidden_scale, freq_emb=self.freq_emb,
                    emb_scale=self.emb_scale, samplerate=self.samplerate, segment=self.segment)


def param_shapes(cfg: HTConfig):
    out = []
    S, C = len(cfg.sources), cfg.audio_channels

    def dconv(prefix, ch):
        hid = int(ch / cfg.dconv_comp)
        for d in range(cfg.dconv_depth):
            p = f"{prefix}.dconv.layers.{d}"
            out.extend([(f"{p}.0.weight", (hid, ch, 3)), (f"{p}.0.bias", (hid,)), (f"{p}.1.weight", (hid,)), (f"{p}.1.bias", (hid,)),
                        (f"{p}.3.weight", (2 * ch, hid, 1)), (f"{p}.3.bias", (2 * ch,)), (f"{p}.4.weight", (2 * ch,)), (f"{p}.4.bias", (2 * ch,)), (f"{p}.6.scale", (ch,))])

    chin, chin_z, chout, chout_z = C, 2 * C, cfg.channels, cfg.channels
    enc, dec, tenc, tdec = [], [], [], []
    for i in range(cfg.depth):
        enc.append((i, chin_z, chout_z))
        tenc.append((i, chin, chout))
        if i == 0:
            chin, chin_z = C * S, 2 * C * S
        dec.insert(0, (chout_z, chin_z))
        tdec.insert(0, (chout, chin))
        chin, chin_z = chout, chout_z
        chout, chout_z = cfg.growth * chout, cfg.growth * chout_z
    K = cfg.kernel_size
    for i, ci, co in enc:
        out.extend([(f"encoder.{i}.conv.weight", (co, ci, K, 1)), (f"encoder.{i}.conv.bias", (co,)), (f"encoder.{i}.rewrite.weight", (2 * co, co, 1, 1)), (f"encoder.{i}.rewrite.bias", (2 * co,))])
        if cfg.dconv_mode & 1:
            dconv(f"encoder.{i}", co)
    k3 = 1 + 2 * cfg.context
    for j, (ci, co) in enumerate(dec):
        out.extend([(f"decoder.{j}.conv_tr.weight", (ci, co, K, 1)), (f"decoder.{j}.conv_tr.bias", (co,)), (f"decoder.{j}.rewrite.weight", (2 * ci, ci, k3, k3)), (f"decoder.{j}.rewrite.bias", (2 * ci,))])
        if cfg.dconv_mode & 2:
            dconv(f"decoder.{j}", ci)
    for i, ci, co in tenc:
        out.extend([(f"tencoder.{i}.conv.weight", (co, ci, K)), (f"tencoder.{i}.conv.bias", (co,)), (f"tencoder.{i}.rewrite.weight", (2 * co, co, 1)), (f"tencoder.{i}.rewrite.bias", (2 * co,))])
        if cfg.dconv_mode & 1:
            dconv(f"tencoder.{i}", co)
    for j, (ci, co) in enumerate(tdec):
        out.extend([(f"tdecoder.{j}.conv_tr.weight", (ci, co, K)), (f"tdecoder.{j}.conv_tr.bias", (co,)), (f"tdecoder.{j}.rewrite.weight", (2 * ci, ci, k3)), (f"tdecoder.{j}.rewrite.bias", (2 * ci,))])
        if cfg.dconv_mode & 2:
            dconv(f"tdecoder.{j}", ci)
    out.append(("freq_emb.embedding.weight", (cfg.nfft // 2 // cfg.stride, cfg.channels)))
    tc = cfg.channels * cfg.growth ** (cfg.depth - 1)
    dim = tc
    if cfg.bottom_channels:
        dim = cfg.bottom_channels
        for nm, (o, i_) in (("channel_upsampler", (dim, tc)), ("channel_downsampler", (tc, dim)), ("channel_upsampler_t", (dim, tc)), ("channel_downsampler_t", (tc, dim))):
            out.extend([(f"{nm}.weight", (o, i_, 1)), (f"{nm}.bias", (o,))])
    hid = int(dim * cfg.t_hidden_scale)
    ct = "crosstransformer"
    for nm in ("norm_in", "norm_in_t"):
        out.extend([(f"{ct}.{nm}.weight", (dim,)), (f"{ct}.{nm}.bias", (dim,))])
    for branch in ("layers", "layers_t"):
        for li in range(cfg.t_layers):
            p = f"{ct}.{branch}.{li}"
            attn = "self_attn" if li % 2 == 0 else "cross_attn"
            out.extend([(f"{p}.{attn}.in_proj_weight", (3 * dim, dim)), (f"{p}.{attn}.in_proj_bias", (3 * dim,)), (f"{p}.{attn}.out_proj.weight", (dim, dim)), (f"{p}.{attn}.out_proj.bias", (dim,)),
                        (f"{p}.linear1.weight", (hid, dim)), (f"{p}.linear1.bias", (hid,)), (f"{p}.linear2.weight", (dim, hid)), (f"{p}.linear2.bias", (dim,))])
            for nm in (("norm1", "norm2") if li % 2 == 0 else ("norm1", "norm2", "norm3")) + ("norm_out",):
                out.extend([(f"{p}.{nm}.weight", (dim,)), (f"{p}.{nm}.bias", (dim,))])
            out.extend([(f"{p}.gamma_1.scale", (dim,)), (f"{p}.gamma_2.scale", (dim,))])
    return out


def make_weights(cfg: HTConfig, seed=0):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in param_shapes(cfg):
        if name.endswith(".scale"):
            a = rng.uniform(0.05, 0.3, shape)
        elif len(shape) == 1 and name.endswith(".weight"):
            a = rng.uniform(0.7, 1.3, shape)
        elif len(shape) == 1 or name.endswith("bias"):
            a = rng.normal(0.0, 0.05, shape)
        elif "freq_emb" in name:
            a = rng.normal(0.0, 0.1, shape)
        elif "conv_tr" in name:
            a = rng.normal(0.0, math.sqrt(2.0 / (shape[0] * 2)), shape)  # each output sees kernel/stride = 2 taps
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.normal(0.0, math.sqrt(1.5 / fan_in), shape)
        w[name] = a.astype(np.float32)
    return w


# ---------------------------------------------------------------------------------------------------------
def sin_embedding_1d(length, dim, shift=0, max_period=10000.0):  # transformer.py:19-26 -> (length, dim)
    pos = shift + np.arange(length, dtype=np.float32)[:, None]
    half = dim // 2
    adim = np.arange(half, dtype=np.float32)[None, :]
    phase = pos / (np.float32(max_period) ** (adim / np.float32(half - 1)))
    return np.concatenate([np.cos(phase), np.sin(phase)], -1).astype(np.float32)


def sin_embedding_2d(d_model, height, width, max_period=10000.0):  # transformer.py:29-49 -> (d_model, height, width)
    import torch

    pe = torch.zeros(d_model, height, width)
    dm = d_model // 2
    div = torch.exp(torch.arange(0.0, dm, 2) * -(math.log(max_period) / dm))
    pw = torch.arange(0.0, width).unsqueeze(1)
    ph = torch.arange(0.0, height).unsqueeze(1)
    pe[0:dm:2] = torch.sin(pw * div).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[1:dm:2] = torch.cos(pw * div).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[dm::2] = torch.sin(ph * div).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    pe[dm + 1 :: 2] = torch.cos(ph * div).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    return pe.numpy()


def forward(weights, cfg: HTConfig, mix: np.ndarray, dtype="float32") -> np.ndarray:
    """HTDemucs.forward in eval mode with use_train_segment=True: mix (B, 2, L <= seg_len) -> (B, S, 2, L)."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items()}
    x_in = torch.from_numpy(np.ascontiguousarray(mix)).to(td)
    S, C = len(cfg.sources), cfg.audio_channels
    hl, nfft = cfg.hop, cfg.nfft
    L0 = x_in.shape[-1]
    T_len = cfg.seg_len
    if L0 < T_len:
        x_in = F.pad(x_in, (0, T_len - L0))  # htdemucs.py:490-493
    mixp = x_in
    B = mixp.shape[0]

    def dconv(y, prefix):  # demucs.py:166-168
        if f"{prefix}.dconv.layers.0.0.weight" not in W:  # dconv_mode without this side
            return y
        for d in range(cfg.dconv_depth):
            p = f"{prefix}.dconv.layers.{d}"
            dil = 2**d
            h = F.conv1d(y, W[f"{p}.0.weight"], W[f"{p}.0.bias"], dilation=dil, padding=dil)
            h = F.gelu(F.group_norm(h, 1, W[f"{p}.1.weight"], W[f"{p}.1.bias"]))
            h = F.conv1d(h, W[f"{p}.3.weight"], W[f"{p}.3.bias"])
            h = F.glu(F.group_norm(h, 1, W[f"{p}.4.weight"], W[f"{p}.4.bias"]), dim=1)
            y = y + W[f"{p}.6.scale"][:, None] * h
        return y

    def enc_layer(x, prefix, freq, inject=None):  # hdemucs.py:119-153
        if freq:
            y = F.conv2d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=(cfg.stride, 1), padding=(cfg.kernel_size // 4, 0))
        else:
            le = x.shape[-1]
            if le % cfg.stride:
                x = F.pad(x, (0, cfg.stride - le % cfg.stride))
            y = F.conv1d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=cfg.stride, padding=cfg.kernel_size // 4)
        if inject is not None:
            y = y + inject
        y = F.gelu(y)
        if freq:
            Bb, Cc, Fr, Tt = y.shape
            y = dconv(y.permute(0, 2, 1, 3).reshape(-1, Cc, Tt), prefix).view(Bb, Fr, Cc, Tt).permute(0, 2, 1, 3)
            z = F.conv2d(y, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"])
        else:
            y = dconv(y, prefix)
            z = F.conv1d(y, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"])
        return F.glu(z, dim=1)

    def dec_layer(x, skip, length, prefix, freq, last):  # hdemucs.py:299-330
        x = x + skip
        pad_c = cfg.context
        if freq:
            y = F.glu(F.conv2d(x, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=pad_c), dim=1)
            Bb, Cc, Fr, Tt = y.shape
            y = dconv(y.permute(0, 2, 1, 3).reshape(-1, Cc, Tt), prefix).view(Bb, Fr, Cc, Tt).permute(0, 2, 1, 3)
            z = F.conv_transpose2d(y, W[f"{prefix}.conv_tr.weight"], W[f"{prefix}.conv_tr.bias"], stride=(cfg.stride, 1))
            pad = cfg.kernel_size // 4
            z = z[..., pad:-pad, :]
        else:
            y = F.glu(F.conv1d(x, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=pad_c), dim=1)
            y = dconv(y, prefix)
            z = F.conv_transpose1d(y, W[f"{prefix}.conv_tr.weight"], W[f"{prefix}.conv_tr.bias"], stride=cfg.stride)
            pad = cfg.kernel_size // 4
            z = z[..., pad : pad + length]
        if not last:
            z = F.gelu(z)
        return z

    def mha(q_in, kv_in, p):  # nn.MultiheadAttention(batch_first=True), eval
        dim = q_in.shape[-1]
        H = cfg.t_heads
        hd = dim // H
        Wi, bi = W[f"{p}.in_proj_weight"], W[f"{p}.in_proj_bias"]
        q = F.linear(q_in, Wi[:dim], bi[:dim])
        k = F.linear(kv_in, Wi[dim : 2 * dim], bi[dim : 2 * dim])
        v = F.linear(kv_in, Wi[2 * dim :], bi[2 * dim :])
        Bq, Lq, _ = q.shape
        Lk = k.shape[1]
        q = q.view(Bq, Lq, H, hd).transpose(1, 2)
        k = k.view(Bq, Lk, H, hd).transpose(1, 2)
        v = v.view(Bq, Lk, H, hd).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(Bq, Lq, dim)
        return F.linear(o, W[f"{p}.out_proj.weight"], W[f"{p}.out_proj.bias"])

    def ln(x, p):
        return F.layer_norm(x, (x.shape[-1],), W[f"{p}.weight"], W[f"{p}.bias"], 1e-5)

    def gn_tokens(x, p):  # MyGroupNorm(1, dim) on (B, T, C): statistics over (T, C) per sample (transformer.py:184-193)
        return F.group_norm(x.transpose(1, 2), 1, W[f"{p}.weight"], W[f"{p}.bias"], 1e-5).transpose(1, 2)

    def ffn(x, p):
        return F.linear(F.gelu(F.linear(x, W[f"{p}.linear1.weight"], W[f"{p}.linear1.bias"])), W[f"{p}.linear2.weight"], W[f"{p}.linear2.bias"])

    def self_layer(x, p):  # transformer.py:268-274
        x = x + W[f"{p}.gamma_1.scale"] * mha(ln(x, f"{p}.norm1"), ln(x, f"{p}.norm1"), f"{p}.self_attn")
        x = x + W[f"{p}.gamma_2.scale"] * ffn(ln(x, f"{p}.norm2"), p)
        return gn_tokens(x, f"{p}.norm_out")

    def cross_layer(q, k, p):  # transformer.py:385-390
        x = q + W[f"{p}.gamma_1.scale"] * mha(ln(q, f"{p}.norm1"), ln(k, f"{p}.norm2"), f"{p}.cross_attn")
        x = x + W[f"{p}.gamma_2.scale"] * ffn(ln(x, f"{p}.norm3"), p)
        return gn_tokens(x, f"{p}.norm_out")

    with torch.no_grad():
        # ---- _spec (:383-403) + _magnitude (:415-424)
        le = int(math.ceil(T_len / hl))
        pad = hl // 2 * 3
        xp = F.pad(mixp, (pad, pad + le * hl - T_len), mode="reflect")
        z = torch.stft(xp.reshape(-1, xp.shape[-1]), nfft, hl, window=torch.hann_window(nfft).to(td), win_length=nfft, normalized=True, center=True, return_complex=True, pad_mode="reflect")
        z = z.view(B, C, z.shape[-2], z.shape[-1])[..., :-1, :][..., 2 : 2 + le]
        Fq, T = z.shape[-2:]
        x = torch.view_as_real(z).permute(0, 1, 4, 2, 3).reshape(B, C * 2, Fq, T)
        mean = x.mean(dim=(1, 2, 3), keepdim=True)
        std = x.std(dim=(1, 2, 3), keepdim=True)
        x = (x - mean) / (1e-5 + std)
        xt = mixp
        meant = xt.mean(dim=(1, 2), keepdim=True)
        stdt = xt.std(dim=(1, 2), keepdim=True)
        xt = (xt - meant) / (1e-5 + stdt)
        saved, saved_t, lengths, lengths_t = [], [], [], []
        for i in range(cfg.depth):  # :520-544
            lengths.append(x.shape[-1])
            lengths_t.append(xt.shape[-1])
            xt = enc_layer(xt, f"tencoder.{i}", False)
            saved_t.append(xt)
            x = enc_layer(x, f"encoder.{i}", True)
            if i == 0:
                emb = (W["freq_emb.embedding.weight"] * cfg.emb_scale).t()[None, :, :, None]  # ScaledEmbedding.forward (hdemucs.py:62-64)
                x = x + cfg.freq_emb * emb
            saved.append(x)
        if cfg.t_layers:
            if cfg.bottom_channels:  # :546-552
                b_, c_, f_, t_ = x.shape
                x = F.conv1d(x.reshape(b_, c_, f_ * t_), W["channel_upsampler.weight"], W["channel_upsampler.bias"]).view(b_, -1, f_, t_)
                xt = F.conv1d(xt, W["channel_upsampler_t.weight"], W["channel_upsampler_t.bias"])
            # CrossTransformerEncoder.forward (transformer.py:529-560)
            Bb, Cc, Fr, T1 = x.shape
            pe2 = torch.from_numpy(sin_embedding_2d(Cc, Fr, T1, cfg.max_period)).to(td)  # (C, Fr, T1)
            pe2 = pe2.permute(2, 1, 0).reshape(1, T1 * Fr, Cc)  # "b c fr t1 -> b (t1 fr) c"
            xs = x.permute(0, 3, 2, 1).reshape(Bb, T1 * Fr, Cc)
            xs = ln(xs, "crosstransformer.norm_in") + pe2
            T2 = xt.shape[-1]
            pe1 = torch.from_numpy(sin_embedding_1d(T2, Cc, 0, cfg.max_period)).to(td)[None]
            xts = ln(xt.permute(0, 2, 1), "crosstransformer.norm_in_t") + pe1
            for li in range(cfg.t_layers):
                if li % 2 == 0:
                    xs = self_layer(xs, f"crosstransformer.layers.{li}")
                    xts = self_layer(xts, f"crosstransformer.layers_t.{li}")
                else:
                    old = xs
                    xs = cross_layer(xs, xts, f"crosstransformer.layers.{li}")
                    xts = cross_layer(xts, old, f"crosstransformer.layers_t.{li}")
            x = xs.reshape(Bb, T1, Fr, Cc).permute(0, 3, 2, 1)
            xt = xts.permute(0, 2, 1)
            if cfg.bottom_channels:  # :556-560
                b_, c_, f_, t_ = x.shape
                x = F.conv1d(x.reshape(b_, c_, f_ * t_), W["channel_downsampler.weight"], W["channel_downsampler.bias"]).view(b_, -1, f_, t_)
                xt = F.conv1d(xt, W["channel_downsampler_t.weight"], W["channel_downsampler_t.bias"])
        for j in range(cfg.depth):  # :562-580
            x = dec_layer(x, saved.pop(-1), lengths.pop(-1), f"decoder.{j}", True, j == cfg.depth - 1)
            xt = dec_layer(xt, saved_t.pop(-1), lengths_t.pop(-1), f"tdecoder.{j}", False, j == cfg.depth - 1)
        x = x.view(B, S, -1, Fq, T) * std[:, None] + mean[:, None]  # :588-589
        zc = torch.view_as_complex(x.view(B, S, -1, 2, Fq, T).permute(0, 1, 2, 4, 5, 3).contiguous())  # _mask (:430-434)
        # ---- _ispec (:405-413)
        zc = F.pad(F.pad(zc, (0, 0, 0, 1)), (2, 2))
        lei = hl * int(math.ceil(T_len / hl)) + 2 * pad
        xi = torch.istft(zc.reshape(-1, zc.shape[-2], zc.shape[-1]), nfft, hl, window=torch.hann_window(nfft).to(td), win_length=nfft, normalized=True, length=lei, center=True)
        xi = xi.view(B, S, C, lei)[..., pad : pad + T_len]
        xt = xt.view(B, S, -1, T_len) * stdt[:, None] + meant[:, None]
        out = xt + xi
        if L0 < T_len:
            out = out[..., :L0]
    return out.to(torch.float32).numpy()


# ---------------------------------------------------------------------------------------------------------
def center_trim(a: np.ndarray, length: int):  # utils.py:53-70
    delta = a.shape[-1] - length
    return a[..., delta // 2 : a.shape[-1] - (delta - delta // 2)] if delta else a


def padded(tensor: np.ndarray, offset: int, length: int, target: int):  # TensorChunk.padded (apply.py:97-113)
    total = tensor.shape[-1]
    delta = target - length
    start = offset - delta // 2
    end = start + target
    cs, ce = max(0, start), min(total, end)
    out = np.zeros(tensor.shape[:-1] + (target,), tensor.dtype)
    out[..., cs - start : cs - start + (ce - cs)] = tensor[..., cs:ce]
    return out


def apply_split(model_fn, cfg: HTConfig, tensor: np.ndarray, offset: int, length: int, overlap=0.25):
    """apply_model(split=True) (apply.py:215-250) on the TensorChunk (tensor, offset, length): -> (1, S, C, length)."""
    S = len(cfg.sources)
    seg = cfg.seg_len
    stride = int((1 - overlap) * seg)
    out = np.zeros((1, S, tensor.shape[1], length), np.float32)
    sw = np.zeros(length, np.float32)
    weight = np.concatenate([np.arange(1, seg // 2 + 1), np.arange(seg - seg // 2, 0, -1)]).astype(np.float32)
    weight = weight / weight.max()
    for off in range(0, length, stride):
        clen = min(length - off, seg)
        chunk = padded(tensor, offset + off, clen, seg)  # leaf: chunk.padded(valid_length) -> model -> center_trim (apply.py:251-260)
        co = center_trim(model_fn(chunk), clen)
        out[..., off : off + seg] += weight[:clen] * co
        sw[off : off + seg] += weight[:clen]
    return out / sw


def apply_model(model_fn, cfg: HTConfig, mix: np.ndarray, shift_offsets, overlap=0.25):
    """apply_model(shifts=len(shift_offsets), split=True) (apply.py:197-214) with the random offsets injected: mix (1,C,N) -> (1,S,C,N)."""
    N = mix.shape[-1]
    if not shift_offsets:
        return apply_split(model_fn, cfg, mix, 0, N, overlap)
    max_shift = int(0.5 * cfg.samplerate)
    pm = padded(mix, 0, N, N + 2 * max_shift)
    out = 0
    for o in shift_offsets:
        so = apply_split(model_fn, cfg, pm, o, N + max_shift - o, overlap)
        out = out + so[..., max_shift - o :]
    return out / len(shift_offsets)


def demix_demucs(model_fns, bag_weights, cfg: HTConfig, mix: np.ndarray, shift_offsets, overlap=0.25):
    """DemucsSeparator.demix_demucs (demucs_separator.py:162-195) with a bag of models (apply.py:169-195):
    mix (2, N) -> sources (S, 2, N) with sources 0 and 1 swapped.  shift_offsets: per model, the injected randint results."""
    import torch

    m = torch.from_numpy(np.asarray(mix, dtype=np.float32))
    ref = m.mean(0)
    mn = ((m - ref.mean()) / ref.std()).numpy()
    S = len(cfg.sources)
    est = np.zeros((1, S, 2, mix.shape[1]), np.float32)
    tot = np.zeros(S, np.float32)
    for fn, w, offs in zip(model_fns, bag_weights, shift_offsets):
        o = apply_model(fn, cfg, mn[None], offs, overlap)
        for k in range(S):
            o[:, k] *= w[k]
            tot[k] += w[k]
        est += o
    est /= tot[None, :, None, None]
    src = (torch.from_numpy(est[0]) * ref.std() + ref.mean()).numpy()
    src[[0, 1]] = src[[1, 0]]
    return src
