"""Host-side parameter plumbing for ConvTDFNet (the network inside UVR-MDX-NET-*.onnx).

The C ABI (b200sep_mdxnet_create) takes ONE flat float32 blob in the reference module's state_dict order
(uvr_lib_v5/mdxnet.py:53-98, modules.py:9-70; `num_batches_tracked`, `window`, `freq_pad` excluded).
"""
from __future__ import annotations

import numpy as np


def convtdfnet_param_names(dim_c, dim_f, num_blocks, l, g, k, bn):
    """Ordered (name, shape) pairs."""
    n = num_blocks // 2
    out = []

    def bn_(prefix, c):
        for nm in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{prefix}.{nm}", (c,)))

    def block(prefix, c, f):
        for i in range(l):
            out.append((f"{prefix}.tfc.H.{i}.0.weight", (c, c, k, k)))
            out.append((f"{prefix}.tfc.H.{i}.0.bias", (c,)))
            bn_(f"{prefix}.tfc.H.{i}.1", c)
        out.append((f"{prefix}.tdf.0.weight", (f // bn, f)))
        bn_(f"{prefix}.tdf.1", c)
        out.append((f"{prefix}.tdf.3.weight", (f, f // bn)))
        bn_(f"{prefix}.tdf.4", c)

    out.append(("first_conv.0.weight", (g, dim_c, 1, 1)))
    out.append(("first_conv.0.bias", (g,)))
    bn_("first_conv.1", g)
    f, c = dim_f, g
    for i in range(n):
        block(f"encoding_blocks.{i}", c, f)
        out.append((f"ds.{i}.0.weight", (c + g, c, 2, 2)))
        out.append((f"ds.{i}.0.bias", (c + g,)))
        bn_(f"ds.{i}.1", c + g)
        f //= 2
        c += g
    block("bottleneck_block", c, f)
    for i in range(n):
        out.append((f"us.{i}.0.weight", (c, c - g, 2, 2)))
        out.append((f"us.{i}.0.bias", (c - g,)))
        bn_(f"us.{i}.1", c - g)
        f *= 2
        c -= g
        block(f"decoding_blocks.{i}", c, f)
    out.append(("final_conv.0.weight", (dim_c, c, 1, 1)))
    out.append(("final_conv.0.bias", (dim_c,)))
    return out


def flatten_state(state: dict, dim_c, dim_f, num_blocks, l, g, k, bn) -> np.ndarray:
    """state: name -> array (numpy or anything np.asarray accepts).  Raises KeyError/ValueError on mismatch."""
    parts = []
    for name, shape in convtdfnet_param_names(dim_c, dim_f, num_blocks, l, g, k, bn):
        a = np.asarray(state[name], dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"parameter {name}: expected shape {shape}, got {a.shape}")
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def infer_hparams_from_state(state: dict):
    """Recover (dim_c, dim_f, num_blocks, l, g, k, bn) from parameter shapes."""
    g, dim_c = state["first_conv.0.weight"].shape[:2]
    n = 0
    while f"encoding_blocks.{n}.tfc.H.0.0.weight" in state:
        n += 1
    l = 0
    while f"encoding_blocks.0.tfc.H.{l}.0.weight" in state:
        l += 1
    k = state["encoding_blocks.0.tfc.H.0.0.weight"].shape[-1]
    f_bn, dim_f = state["encoding_blocks.0.tdf.0.weight"].shape
    return dict(dim_c=int(dim_c), dim_f=int(dim_f), num_blocks=2 * n + 1, l=l, g=int(g), k=int(k), bn=int(dim_f // f_bn))
