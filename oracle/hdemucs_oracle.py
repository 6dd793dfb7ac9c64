"""CPU oracle for Hybrid Demucs v3 (HDemucs: conv + BiLSTM + LocalState DConv branches) -- TEST INFRASTRUCTURE, not product code.

Functional torch-CPU restatement of
  * HDemucs.__init__ layer plan / forward   uvr_lib_v5/demucs/hdemucs.py:360-527 (plan), :665-783 (forward), _spec :529-548, _ispec :550-568,
                                            _magnitude :570-579, _mask (cac branch) :581-590
  * HEncLayer / HDecLayer                   uvr_lib_v5/demucs/hdemucs.py:67-153, :252-330 (GroupNorm layers, "empty" merge layers, inject, last_freq)
  * DConv with BLSTM and LocalState         uvr_lib_v5/demucs/demucs.py:19-67 (BLSTM, overlapping 200-step frames), :99-168 (DConv), :171-231 (LocalState)
for the structure of the released hybrid models (hdemucs_mmi, mdx_extra*): hybrid, complex-as-channels, no multi_freqs, rewrite, no Wiener
filtering (cac).  apply_model / demix_demucs are the ones of demucs_oracle.py with the model's own segment and NO valid_length padding
(HDemucs has no valid_length, apply.py:252-257: the last segment runs at its own length).
Pinned against the unmodified reference by oracle/make_golden_hdemucs.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class HDConfig:
    sources: tuple = ("drums", "bass", "other", "vocals")
    audio_channels: int = 2
    channels: int = 48
    growth: int = 2
    nfft: int = 4096
    depth: int = 6
    hybrid_old: bool = False
    freq_emb: float = 0.2
    emb_scale: float = 10.0
    kernel_size: int = 8
    time_stride: int = 2
    stride: int = 4
    context: int = 1
    context_enc: int = 0
    norm_starts: int = 4
    norm_groups: int = 4
    dconv_mode: int = 1
    dconv_depth: int = 2
    dconv_comp: int = 4
    dconv_attn: int = 4
    dconv_lstm: int = 4
    samplerate: int = 44100
    segment: float = 40.0
    # LocalState / BLSTM constants of DConv (demucs.py:99, :152-155)
    attn_heads: int = 4
    attn_ndecay: int = 4
    lstm_layers: int = 2
    lstm_max_steps: int = 200

    @property
    def hop(self):
        return self.nfft // 4

    @property
    def seg_len(self):  # apply.py:218
        return int(self.samplerate * self.segment)

    def kwargs(self):  # what the reference constructor receives
        return dict(sources=list(self.sources), audio_channels=self.audio_channels, channels=self.channels, growth=self.growth, nfft=self.nfft, depth=self.depth,
                    hybrid_old=self.hybrid_old, freq_emb=self.freq_emb, emb_scale=self.emb_scale, kernel_size=self.kernel_size, time_stride=self.time_stride,
                    stride=self.stride, context=self.context, context_enc=self.context_enc, norm_starts=self.norm_starts, norm_groups=self.norm_groups,
                    dconv_mode=self.dconv_mode, dconv_depth=self.dconv_depth, dconv_comp=self.dconv_comp, dconv_attn=self.dconv_attn, dconv_lstm=self.dconv_lstm,
                    samplerate=self.samplerate, segment=self.segment)


def layer_plan(cfg: HDConfig):
    """The per-index layer geometry HDemucs.__init__ derives (hdemucs.py:455-527).  One dict per index with the encoder ("enc"), time
    encoder ("tenc", or None), decoder ("dec") and time decoder ("tdec", or None) descriptions."""
    S, C = len(cfg.sources), cfg.audio_channels
    chin, chin_z = C, 2 * C
    chout, chout_z = cfg.channels, cfg.channels
    freqs = cfg.nfft // 2
    plan = []
    for index in range(cfg.depth):
        lstm, attn, norm = index >= cfg.dconv_lstm, index >= cfg.dconv_attn, index >= cfg.norm_starts
        freq = freqs > 1
        stri, ker = cfg.stride, cfg.kernel_size
        if not freq:
            assert freqs == 1
            ker, stri = cfg.time_stride * 2, cfg.time_stride
        pad, last_freq = True, False
        if freq and freqs <= cfg.kernel_size:
            ker, pad, last_freq = freqs, False, True
        if last_freq:
            chout_z = max(chout, chout_z)
            chout = chout_z
        common = dict(norm=norm, lstm=lstm, attn=attn)
        e = dict(common, chin=chin_z, chout=chout_z, k=ker, s=stri, freq=freq, pad=(ker // 4 if pad else 0), dconv=bool(cfg.dconv_mode & 1), context=cfg.context_enc, empty=False)
        te = None
        if freq:
            te = dict(common, chin=chin, chout=chout, k=cfg.kernel_size, s=cfg.stride, freq=False, pad=cfg.kernel_size // 4, dconv=bool(cfg.dconv_mode & 1), context=cfg.context_enc,
                      empty=last_freq)
        if index == 0:
            chin = C * S
            chin_z = 2 * chin
        d = dict(common, chin=chout_z, chout=chin_z, k=ker, s=stri, freq=freq, pad=(ker // 4 if pad else 0), dconv=bool(cfg.dconv_mode & 2), context=cfg.context, empty=False,
                 last=index == 0)
        td = None
        if freq:
            td = dict(common, chin=chout, chout=chin, k=cfg.kernel_size, s=cfg.stride, freq=False, pad=cfg.kernel_size // 4, dconv=bool(cfg.dconv_mode & 2), context=cfg.context,
                      empty=last_freq, last=index == 0)
        plan.append(dict(enc=e, tenc=te, dec=d, tdec=td))
        chin, chin_z = chout, chout_z
        chout, chout_z = int(cfg.growth * chout), int(cfg.growth * chout_z)
        if freq:
            freqs = 1 if freqs <= cfg.kernel_size else freqs // cfg.stride
    return plan


def param_shapes(cfg: HDConfig):
    """state_dict names and shapes in the reference's registration order (encoder, decoder, tencoder, tdecoder interleaved per index as the
    constructor appends / inserts them is NOT the state_dict order: ModuleLists are walked one after the other)."""
    plan = layer_plan(cfg)
    out = []

    def dconv(prefix, ch, lstm, attn):
        hid = int(ch / cfg.dconv_comp)
        for d in range(cfg.dconv_depth):
            p = f"{prefix}.dconv.layers.{d}"
            out.extend([(f"{p}.0.weight", (hid, ch, 3)), (f"{p}.0.bias", (hid,)), (f"{p}.1.weight", (hid,)), (f"{p}.1.bias", (hid,))])
            i = 3
            if lstm:
                for layer in range(cfg.lstm_layers):
                    for sfx in ("", "_reverse"):
                        nin = hid if layer == 0 else 2 * hid
                        out.extend([(f"{p}.{i}.lstm.weight_ih_l{layer}{sfx}", (4 * hid, nin)), (f"{p}.{i}.lstm.weight_hh_l{layer}{sfx}", (4 * hid, hid)),
                                    (f"{p}.{i}.lstm.bias_ih_l{layer}{sfx}", (4 * hid,)), (f"{p}.{i}.lstm.bias_hh_l{layer}{sfx}", (4 * hid,))])
                out.extend([(f"{p}.{i}.linear.weight", (hid, 2 * hid)), (f"{p}.{i}.linear.bias", (hid,))])
                i += 1
            if attn:
                for nm, co in (("content", hid), ("query", hid), ("key", hid), ("query_decay", cfg.attn_heads * cfg.attn_ndecay), ("proj", hid)):
                    out.extend([(f"{p}.{i}.{nm}.weight", (co, hid, 1)), (f"{p}.{i}.{nm}.bias", (co,))])
                i += 1
            out.extend([(f"{p}.{i}.weight", (2 * ch, hid, 1)), (f"{p}.{i}.bias", (2 * ch,)), (f"{p}.{i + 1}.weight", (2 * ch,)), (f"{p}.{i + 1}.bias", (2 * ch,)),
                        (f"{p}.{i + 3}.scale", (ch,))])

    def enc(prefix, L):
        kshape = (L["k"], 1) if L["freq"] else (L["k"],)
        one = (1 + 2 * L["context"],) * (2 if L["freq"] else 1)
        out.extend([(f"{prefix}.conv.weight", (L["chout"], L["chin"]) + kshape), (f"{prefix}.conv.bias", (L["chout"],))])
        if L["empty"]:
            return
        if L["norm"]:
            out.extend([(f"{prefix}.norm1.weight", (L["chout"],)), (f"{prefix}.norm1.bias", (L["chout"],))])
        out.extend([(f"{prefix}.rewrite.weight", (2 * L["chout"], L["chout"]) + one), (f"{prefix}.rewrite.bias", (2 * L["chout"],))])
        if L["norm"]:
            out.extend([(f"{prefix}.norm2.weight", (2 * L["chout"],)), (f"{prefix}.norm2.bias", (2 * L["chout"],))])
        if L["dconv"]:
            dconv(prefix, L["chout"], L["lstm"], L["attn"])

    def dec(prefix, L):
        kshape = (L["k"], 1) if L["freq"] else (L["k"],)
        one = (1 + 2 * L["context"],) * (2 if L["freq"] else 1)
        out.extend([(f"{prefix}.conv_tr.weight", (L["chin"], L["chout"]) + kshape), (f"{prefix}.conv_tr.bias", (L["chout"],))])
        if L["norm"]:
            out.extend([(f"{prefix}.norm2.weight", (L["chout"],)), (f"{prefix}.norm2.bias", (L["chout"],))])
        if L["empty"]:
            return
        out.extend([(f"{prefix}.rewrite.weight", (2 * L["chin"], L["chin"]) + one), (f"{prefix}.rewrite.bias", (2 * L["chin"],))])
        if L["norm"]:
            out.extend([(f"{prefix}.norm1.weight", (2 * L["chin"],)), (f"{prefix}.norm1.bias", (2 * L["chin"],))])
        if L["dconv"]:
            dconv(prefix, L["chin"], L["lstm"], L["attn"])

    for i, P in enumerate(plan):
        enc(f"encoder.{i}", P["enc"])
    decs = [P["dec"] for P in plan][::-1]
    for j, L in enumerate(decs):
        dec(f"decoder.{j}", L)
    tencs = [P["tenc"] for P in plan if P["tenc"] is not None]
    for i, L in enumerate(tencs):
        enc(f"tencoder.{i}", L)
    tdecs = [P["tdec"] for P in plan if P["tdec"] is not None][::-1]
    for j, L in enumerate(tdecs):
        dec(f"tdecoder.{j}", L)
    if cfg.freq_emb:
        out.append(("freq_emb.embedding.weight", (cfg.nfft // 2 // cfg.stride, plan[0]["enc"]["chout"])))
    return out


def make_weights(cfg: HDConfig, seed=0):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in param_shapes(cfg):
        if name.endswith(".scale"):
            a = rng.uniform(0.05, 0.3, shape)
        elif ".lstm." in name:
            hid = shape[0] // 4
            a = rng.uniform(-1.0, 1.0, shape) / math.sqrt(hid)
        elif "query_decay.bias" in name:
            a = rng.normal(-2.0, 0.3, shape)
        elif "query_decay.weight" in name:
            a = rng.normal(0.0, 0.3 / math.sqrt(shape[1]), shape)
        elif len(shape) == 1 and name.endswith(".weight"):
            a = rng.uniform(0.7, 1.3, shape)
        elif len(shape) == 1 or name.endswith("bias"):
            a = rng.normal(0.0, 0.05, shape)
        elif "freq_emb" in name:
            a = rng.normal(0.0, 0.1, shape)
        elif "conv_tr" in name:
            a = rng.normal(0.0, math.sqrt(2.0 / (shape[0] * 2)), shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.normal(0.0, math.sqrt(1.5 / fan_in), shape)
        w[name] = a.astype(np.float32)
    return w


# ---------------------------------------------------------------------------------------------------------
def lstm_bidir(x, W, p, layers):
    """nn.LSTM(bidirectional=True, num_layers=layers) on x (T, N, hid), zero initial state: gate order i, f, g, o (torch convention)."""
    import torch

    for layer in range(layers):
        outs = []
        for sfx in ("", "_reverse"):
            wih, whh = W[f"{p}.weight_ih_l{layer}{sfx}"], W[f"{p}.weight_hh_l{layer}{sfx}"]
            b = W[f"{p}.bias_ih_l{layer}{sfx}"] + W[f"{p}.bias_hh_l{layer}{sfx}"]
            hid = whh.shape[1]
            xp = x @ wih.t() + b
            h = torch.zeros(x.shape[1], hid, dtype=x.dtype)
            c = torch.zeros_like(h)
            ys = [None] * x.shape[0]
            order = range(x.shape[0] - 1, -1, -1) if sfx else range(x.shape[0])
            for t in order:
                g = xp[t] + h @ whh.t()
                i_, f_, g_, o_ = g[:, :hid], g[:, hid : 2 * hid], g[:, 2 * hid : 3 * hid], g[:, 3 * hid :]
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                h = torch.sigmoid(o_) * torch.tanh(c)
                ys[t] = h
            outs.append(torch.stack(ys))
        x = torch.cat(outs, -1)
    return x


def blstm(x, W, p, cfg: HDConfig):
    """BLSTM(dim, layers=2, max_steps=200, skip=True).forward (demucs.py:34-67) on x (B, C, T)."""
    import torch

    B, C, T = x.shape
    y = x
    framed = False
    width = cfg.lstm_max_steps
    if width is not None and T > width:
        stride = width // 2
        # unfold (utils.py): frames of `width` every `stride`, the signal zero-padded on the right to a whole number of frames
        nframes = int(math.ceil(T / stride))
        tgt = (nframes - 1) * stride + width
        xp = torch.nn.functional.pad(x, (0, tgt - T))
        frames = xp.unfold(-1, width, stride)  # (B, C, nframes, width)
        assert frames.shape[2] == nframes
        framed = True
        x = frames.permute(0, 2, 1, 3).reshape(-1, C, width)
    x = x.permute(2, 0, 1)
    x = lstm_bidir(x, W, f"{p}.lstm", cfg.lstm_layers)
    x = x @ W[f"{p}.linear.weight"].t() + W[f"{p}.linear.bias"]
    x = x.permute(1, 2, 0)
    if framed:
        fr = x.reshape(B, -1, C, width)
        limit = stride // 2
        out = []
        for k in range(nframes):
            if k == 0:
                out.append(fr[:, k, :, :-limit])
            elif k == nframes - 1:
                out.append(fr[:, k, :, limit:])
            else:
                out.append(fr[:, k, :, limit:-limit])
        x = torch.cat(out, -1)[..., :T]
    return x + y


def local_state(x, W, p, cfg: HDConfig):
    """LocalState(channels, heads=4, nfreqs=0, ndecay=4).forward (demucs.py:197-231) on x (B, C, T)."""
    import torch
    import torch.nn.functional as F

    B, C, T = x.shape
    H, nd = cfg.attn_heads, cfg.attn_ndecay
    idx = torch.arange(T, dtype=x.dtype)
    delta = idx[:, None] - idx[None, :]
    q = F.conv1d(x, W[f"{p}.query.weight"], W[f"{p}.query.bias"]).view(B, H, -1, T)
    k = F.conv1d(x, W[f"{p}.key.weight"], W[f"{p}.key.bias"]).view(B, H, -1, T)
    dots = torch.einsum("bhct,bhcs->bhts", k, q) / k.shape[2] ** 0.5
    decays = torch.arange(1, nd + 1, dtype=x.dtype)
    dq = torch.sigmoid(F.conv1d(x, W[f"{p}.query_decay.weight"], W[f"{p}.query_decay.bias"]).view(B, H, -1, T)) / 2
    dk = -decays.view(-1, 1, 1) * delta.abs() / nd**0.5
    dots = dots + torch.einsum("fts,bhfs->bhts", dk, dq)
    dots = dots.masked_fill(torch.eye(T, dtype=torch.bool), -100)
    wts = torch.softmax(dots, dim=2)
    content = F.conv1d(x, W[f"{p}.content.weight"], W[f"{p}.content.bias"]).view(B, H, -1, T)
    res = torch.einsum("bhts,bhct->bhcs", wts, content).reshape(B, -1, T)
    return x + F.conv1d(res, W[f"{p}.proj.weight"], W[f"{p}.proj.bias"])


def forward(weights, cfg: HDConfig, mix: np.ndarray, dtype="float32", taps=None) -> np.ndarray:
    """HDemucs.forward in eval mode: mix (B, 2, L) -> (B, S, 2, L).  `taps` (a dict) receives named intermediates for layer-by-layer debugging."""
    import torch
    import torch.nn.functional as F

    td = torch.float64 if dtype == "float64" else torch.float32
    W = {k: torch.from_numpy(np.asarray(v)).to(td) for k, v in weights.items()}
    mixp = torch.from_numpy(np.ascontiguousarray(mix)).to(td)
    S, C = len(cfg.sources), cfg.audio_channels
    hl, nfft = cfg.hop, cfg.nfft
    B, _, length = mixp.shape
    plan = layer_plan(cfg)
    G = cfg.norm_groups

    def tap(name, t):
        if taps is not None:
            taps[name] = t.to(torch.float32).numpy().copy()

    def norm(x, p, on):
        return F.group_norm(x, G, W[f"{p}.weight"], W[f"{p}.bias"], 1e-5) if on else x

    def dconv(y, prefix, L):  # demucs.py:152-168, y (N, C, T)
        for d in range(cfg.dconv_depth):
            p = f"{prefix}.dconv.layers.{d}"
            dil = 2**d
            h = F.conv1d(y, W[f"{p}.0.weight"], W[f"{p}.0.bias"], dilation=dil, padding=dil)
            h = F.gelu(F.group_norm(h, 1, W[f"{p}.1.weight"], W[f"{p}.1.bias"]))
            i = 3
            if L["lstm"]:
                h = blstm(h, W, f"{p}.{i}", cfg)
                i += 1
            if L["attn"]:
                h = local_state(h, W, f"{p}.{i}", cfg)
                i += 1
            h = F.conv1d(h, W[f"{p}.{i}.weight"], W[f"{p}.{i}.bias"])
            h = F.glu(F.group_norm(h, 1, W[f"{p}.{i + 1}.weight"], W[f"{p}.{i + 1}.bias"]), dim=1)
            y = y + W[f"{p}.{i + 3}.scale"][:, None] * h
        return y

    def enc_layer(x, prefix, L, inject=None):  # hdemucs.py:119-153
        if not L["freq"] and x.dim() == 4:
            Bb, Cc, Fr, Tt = x.shape
            x = x.reshape(Bb, -1, Tt)
        if L["freq"]:
            y = F.conv2d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=(L["s"], 1), padding=(L["pad"], 0))
        else:
            le = x.shape[-1]
            if le % L["s"]:
                x = F.pad(x, (0, L["s"] - le % L["s"]))
            y = F.conv1d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=L["s"], padding=L["pad"])
        if L["empty"]:
            return y
        if inject is not None:
            if inject.dim() == 3 and y.dim() == 4:
                inject = inject[:, :, None]
            y = y + inject
        y = F.gelu(norm(y, f"{prefix}.norm1", L["norm"]))
        if L["dconv"]:
            if L["freq"]:
                Bb, Cc, Fr, Tt = y.shape
                y = dconv(y.permute(0, 2, 1, 3).reshape(-1, Cc, Tt), prefix, L).view(Bb, Fr, Cc, Tt).permute(0, 2, 1, 3)
            else:
                y = dconv(y, prefix, L)
        ctx = L["context"]
        if L["freq"]:
            z = F.conv2d(y, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=ctx)
        else:
            z = F.conv1d(y, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=ctx)
        return F.glu(norm(z, f"{prefix}.norm2", L["norm"]), dim=1)

    def dec_layer(x, skip, length_, prefix, L):  # hdemucs.py:299-330 -> (z, y)
        if L["freq"] and x.dim() == 3:
            Bb, Cc, Tt = x.shape
            x = x.view(Bb, L["chin"], -1, Tt)
        if not L["empty"]:
            x = x + skip
            ctx = L["context"]
            if L["freq"]:
                y = F.conv2d(x, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=ctx)
            else:
                y = F.conv1d(x, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=ctx)
            y = F.glu(norm(y, f"{prefix}.norm1", L["norm"]), dim=1)
            if L["dconv"]:
                if L["freq"]:
                    Bb, Cc, Fr, Tt = y.shape
                    y = dconv(y.permute(0, 2, 1, 3).reshape(-1, Cc, Tt), prefix, L).view(Bb, Fr, Cc, Tt).permute(0, 2, 1, 3)
                else:
                    y = dconv(y, prefix, L)
        else:
            y = x
        if L["freq"]:
            z = F.conv_transpose2d(y, W[f"{prefix}.conv_tr.weight"], W[f"{prefix}.conv_tr.bias"], stride=(L["s"], 1))
        else:
            z = F.conv_transpose1d(y, W[f"{prefix}.conv_tr.weight"], W[f"{prefix}.conv_tr.bias"], stride=L["s"])
        z = norm(z, f"{prefix}.norm2", L["norm"])
        if L["freq"]:
            if L["pad"]:
                z = z[..., L["pad"] : -L["pad"], :]
        else:
            z = z[..., L["pad"] : L["pad"] + length_]
        if not L["last"]:
            z = F.gelu(z)
        return z, y

    with torch.no_grad():
        # ---- _spec (:529-548) + _magnitude (:570-579)
        le = int(math.ceil(length / hl))
        pad = hl // 2 * 3
        if not cfg.hybrid_old:
            xp = F.pad(mixp, (pad, pad + le * hl - length), mode="reflect")
        else:
            xp = F.pad(mixp, (pad, pad + le * hl - length))
        z = torch.stft(xp.reshape(-1, xp.shape[-1]), nfft, hl, window=torch.hann_window(nfft).to(td), win_length=nfft, normalized=True, center=True, return_complex=True, pad_mode="reflect")
        z = z.view(B, C, z.shape[-2], z.shape[-1])[..., :-1, :]
        assert z.shape[-1] == le + 4
        z = z[..., 2 : 2 + le]
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
        tencs = [P["tenc"] for P in plan if P["tenc"] is not None]
        tdecs = [P["tdec"] for P in plan if P["tdec"] is not None][::-1]
        decs = [P["dec"] for P in plan][::-1]
        for idx, P in enumerate(plan):  # :697-716
            lengths.append(x.shape[-1])
            inject = None
            if idx < len(tencs):
                lengths_t.append(xt.shape[-1])
                xt = enc_layer(xt, f"tencoder.{idx}", tencs[idx])
                tap(f"tenc{idx}", xt)
                if not tencs[idx]["empty"]:
                    saved_t.append(xt)
                else:
                    inject = xt
            x = enc_layer(x, f"encoder.{idx}", P["enc"], inject)
            if idx == 0 and cfg.freq_emb:
                emb = (W["freq_emb.embedding.weight"] * cfg.emb_scale).t()[None, :, :, None]  # ScaledEmbedding.forward (hdemucs.py:62-64)
                x = x + cfg.freq_emb * emb
            tap(f"enc{idx}", x)
            saved.append(x)
        x = torch.zeros_like(x)
        xt = torch.zeros_like(x)
        offset = cfg.depth - len(tdecs)
        for idx, L in enumerate(decs):  # :724-749
            skip = saved.pop(-1)
            x, pre = dec_layer(x, skip, lengths.pop(-1), f"decoder.{idx}", L)
            tap(f"dec{idx}", x)
            if idx >= offset:
                Lt = tdecs[idx - offset]
                length_t = lengths_t.pop(-1)
                if Lt["empty"]:
                    assert pre.shape[2] == 1, pre.shape
                    xt, _ = dec_layer(pre[:, :, 0], None, length_t, f"tdecoder.{idx - offset}", Lt)
                else:
                    xt, _ = dec_layer(xt, saved_t.pop(-1), length_t, f"tdecoder.{idx - offset}", Lt)
                tap(f"tdec{idx - offset}", xt)
        assert not saved and not saved_t and not lengths_t
        x = x.view(B, S, -1, Fq, T) * std[:, None] + mean[:, None]
        zc = torch.view_as_complex(x.view(B, S, -1, 2, Fq, T).permute(0, 1, 2, 4, 5, 3).contiguous())  # _mask, cac (:581-590)
        # ---- _ispec (:550-568)
        zc = F.pad(F.pad(zc, (0, 0, 0, 1)), (2, 2))
        if not cfg.hybrid_old:
            lei = hl * int(math.ceil(length / hl)) + 2 * pad
        else:
            lei = hl * int(math.ceil(length / hl))
        xi = torch.istft(zc.reshape(-1, zc.shape[-2], zc.shape[-1]), nfft, hl, window=torch.hann_window(nfft).to(td), win_length=nfft, normalized=True, length=lei, center=True)
        xi = xi.view(B, S, C, lei)
        xi = xi[..., pad : pad + length] if not cfg.hybrid_old else xi[..., :length]
        xt = xt.view(B, S, -1, length) * stdt[:, None] + meant[:, None]
        out = xt + xi
    return out.to(torch.float32).numpy()


# ---------------------------------------------------------------------------------------------------------
def apply_split(model_fn, cfg: HDConfig, tensor: np.ndarray, offset: int, length: int, overlap=0.25):
    """apply_model(split=True) (apply.py:215-250) for a model WITHOUT valid_length: each TensorChunk runs at its own length
    (the leaf call pads to valid_length = length, apply.py:252-260)."""
    from demucs_oracle import padded

    S = len(cfg.sources)
    seg = cfg.seg_len
    stride = int((1 - overlap) * seg)
    out = np.zeros((1, S, tensor.shape[1], length), np.float32)
    sw = np.zeros(length, np.float32)
    weight = np.concatenate([np.arange(1, seg // 2 + 1), np.arange(seg - seg // 2, 0, -1)]).astype(np.float32)
    weight = weight / weight.max()
    for off in range(0, length, stride):
        clen = min(length - off, seg)
        co = model_fn(padded(tensor, offset + off, clen, clen))
        out[..., off : off + seg] += weight[:clen] * co
        sw[off : off + seg] += weight[:clen]
    return out / sw


def apply_model(model_fn, cfg: HDConfig, mix: np.ndarray, shift_offsets, overlap=0.25):
    """apply_model(shifts=len(shift_offsets), split=True) (apply.py:197-214), offsets injected: mix (1,C,N) -> (1,S,C,N)."""
    from demucs_oracle import padded

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
