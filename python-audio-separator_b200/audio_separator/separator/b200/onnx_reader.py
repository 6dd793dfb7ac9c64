"""Minimal ONNX reader for the UVR-MDX-NET-*.onnx graphs (no `onnx` / `onnxruntime` wheel needed).

Replaces `ort.InferenceSession(model_path)` / `onnx2torch.convert` at load time
(audio_separator/separator/architectures/mdx_separator.py:115-131): a protobuf *wire-format* parser pulls the
graph's nodes and initialisers, and a structural walk maps them onto the ConvTDFNet parameter names
(uvr_lib_v5/mdxnet.py:53-98).  The walk relies on execution order only, not on initialiser names, so it accepts
both exports that keep BatchNormalization nodes and exports where Conv+BN were fused by the exporter.

ONNX field numbers used (onnx.proto3): ModelProto.graph=7; GraphProto.node=1, .initializer=5;
NodeProto.input=1, .output=2, .name=3, .op_type=4, .attribute=5; AttributeProto.name=1, .t=5, .ints=8, .i=3;
TensorProto.dims=1, .data_type=2, .float_data=4, .int64_data=7, .name=8, .raw_data=9.
"""
from __future__ import annotations

import struct

import numpy as np


def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message; value is int or memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos : pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos : pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos : pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _packed_varints(v):
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


_DTYPES = {1: np.float32, 7: np.int64, 11: np.float64, 6: np.int32, 10: np.float16}


def _tensor(buf):
    dims, dtype, name, raw, floats, int64s = [], 1, "", None, [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims.extend(_packed_varints(v) if wt == 2 else [v])
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", bytes(v)) if wt == 2 else struct.unpack("<f", bytes(v)))
        elif fno == 7:
            int64s.extend(_packed_varints(v) if wt == 2 else [v])
    if dtype not in _DTYPES:
        raise ValueError(f"initializer {name}: unsupported ONNX data_type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dtype]).newbyteorder("<")).astype(_DTYPES[dtype])
    elif floats:
        arr = np.asarray(floats, dtype=np.float32)
    else:
        arr = np.asarray(int64s, dtype=np.int64)
    return name, arr.reshape(dims) if dims else arr


def _node(buf):
    node = {"input": [], "output": [], "name": "", "op": "", "attrs": {}}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            node["input"].append(bytes(v).decode())
        elif fno == 2:
            node["output"].append(bytes(v).decode())
        elif fno == 3:
            node["name"] = bytes(v).decode()
        elif fno == 4:
            node["op"] = bytes(v).decode()
        elif fno == 5:
            aname, aval = "", None
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    aname = bytes(v2).decode()
                elif f2 == 5:
                    aval = _tensor(v2)[1]
                elif f2 == 8:
                    aval = (aval or []) + (_packed_varints(v2) if w2 == 2 else [v2])
                elif f2 == 3:
                    aval = v2
            node["attrs"][aname] = aval
    return node


def read_graph(path):
    """-> (nodes in execution order, {initializer name: ndarray})"""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    graph = None
    for fno, wt, v in _fields(buf):
        if fno == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found (not an ONNX model?)")
    nodes, inits = [], {}
    for fno, wt, v in _fields(graph):
        if fno == 1:
            nodes.append(_node(v))
        elif fno == 5:
            name, arr = _tensor(v)
            inits[name] = arr
    for nd in nodes:  # Constant nodes act as initialisers
        if nd["op"] == "Constant" and "value" in nd["attrs"] and nd["output"]:
            inits[nd["output"][0]] = nd["attrs"]["value"]
    return nodes, inits


def load_convtdfnet_state(path) -> dict:
    """Map the graph onto ConvTDFNet state_dict names by walking Conv / ConvTranspose / MatMul /
    BatchNormalization nodes in execution order."""
    nodes, inits = read_graph(path)
    consumers = {}
    for nd in nodes:
        for i in nd["input"]:
            consumers.setdefault(i, []).append(nd)

    def following_bn(nd):
        nxt = consumers.get(nd["output"][0], [])
        return nxt[0] if len(nxt) == 1 and nxt[0]["op"] == "BatchNormalization" else None

    layers = []  # (kind, weight, bias|None, bn(4)|None)
    for nd in nodes:
        if nd["op"] in ("Conv", "ConvTranspose"):
            w = inits[nd["input"][1]]
            b = inits[nd["input"][2]] if len(nd["input"]) > 2 and nd["input"][2] else None
            kind = "convT" if nd["op"] == "ConvTranspose" else "conv"
        elif nd["op"] == "MatMul":
            wname = nd["input"][1] if nd["input"][1] in inits else nd["input"][0]
            w = np.ascontiguousarray(inits[wname].T)  # exported as x @ W^T -> initializer is (in, out)
            b, kind = None, "linear"
        elif nd["op"] == "Gemm":
            w = inits[nd["input"][1]]
            if not nd["attrs"].get("transB", 0):
                w = np.ascontiguousarray(w.T)
            b, kind = None, "linear"
        else:
            continue
        bn = following_bn(nd)
        bnp = tuple(inits[n] for n in bn["input"][1:5]) if bn is not None else None
        layers.append((kind, np.asarray(w, dtype=np.float32), b, bnp))

    state = {}
    it = iter(layers)

    def put(prefix_conv, prefix_bn, kind, with_bias=True, need_bn=True):
        k, w, b, bn = next(it)
        if k != kind:
            raise ValueError(f"{path}: expected a {kind} layer for {prefix_conv}, found {k} (not a ConvTDFNet graph?)")
        state[f"{prefix_conv}.weight"] = w
        c_out = w.shape[1] if kind == "convT" else w.shape[0]
        if with_bias:
            state[f"{prefix_conv}.bias"] = np.zeros(c_out, np.float32) if b is None else np.asarray(b, np.float32)
        if need_bn:
            n_bn = c_out if kind != "linear" else None
            if bn is None:
                if kind == "linear":
                    raise ValueError(f"{path}: TDF linear {prefix_conv} is not followed by BatchNormalization")
                # exporter fused Conv+BN: identity statistics so that folding reproduces the fused weights
                bn = (np.ones(n_bn, np.float32), np.zeros(n_bn, np.float32), np.zeros(n_bn, np.float32), np.full(n_bn, 1.0 - 1e-5, np.float32))
            for nm, a in zip(("weight", "bias", "running_mean", "running_var"), bn):
                state[f"{prefix_bn}.{nm}"] = np.asarray(a, np.float32)

    n_convT = sum(1 for l_ in layers if l_[0] == "convT")
    n_lin = sum(1 for l_ in layers if l_[0] == "linear")
    n_blocks = 2 * n_convT + 1
    if n_lin != 2 * n_blocks:
        raise ValueError(f"{path}: {n_lin} linear layers for {n_blocks} TFC-TDF blocks; unsupported topology (bn=0 or no TDF)")
    n_conv = sum(1 for l_ in layers if l_[0] == "conv")
    l = (n_conv - 2 - n_convT) // n_blocks

    def block(prefix):
        for i in range(l):
            put(f"{prefix}.tfc.H.{i}.0", f"{prefix}.tfc.H.{i}.1", "conv")
        put(f"{prefix}.tdf.0", f"{prefix}.tdf.1", "linear", with_bias=False)
        put(f"{prefix}.tdf.3", f"{prefix}.tdf.4", "linear", with_bias=False)

    put("first_conv.0", "first_conv.1", "conv")
    for i in range(n_convT):
        block(f"encoding_blocks.{i}")
        put(f"ds.{i}.0", f"ds.{i}.1", "conv")
    block("bottleneck_block")
    for i in range(n_convT):
        put(f"us.{i}.0", f"us.{i}.1", "convT")
        block(f"decoding_blocks.{i}")
    put("final_conv.0", None, "conv", need_bn=False)
    if next(it, None) is not None:
        raise ValueError(f"{path}: unexpected extra layers after final_conv")
    return state
