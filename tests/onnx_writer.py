"""Test helper: serialise a ConvTDFNet state dict as an ONNX ModelProto (protobuf wire format by hand), in the
node order a torch.onnx export of uvr_lib_v5/mdxnet.py produces.  fuse_conv_bn=True emulates exporters that fold
BatchNorm into the preceding Conv."""
import struct

import numpy as np


def _varint(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(fno, payload):
    return _varint((fno << 3) | 2) + _varint(len(payload)) + payload


def _vi(fno, x):
    return _varint(fno << 3) + _varint(x)


def _tensor(name, arr):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    body = b"".join(_vi(1, d) for d in arr.shape) + _vi(2, 1) + _ld(8, name.encode()) + _ld(9, arr.tobytes())
    return body


def _node(op, inputs, outputs, name=""):
    body = b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs)
    return body + _ld(3, name.encode()) + _ld(4, op.encode())


def write_convtdfnet_onnx(path, state, num_blocks, l, fuse_conv_bn=False, raw_names=False):
    nodes, inits = [], []
    counter = [0]
    cur = ["input"]

    def tname(n):
        counter[0] += 1
        return f"onnx::W_{counter[0]}" if raw_names else n

    def add_init(name, arr):
        nm = tname(name)
        inits.append(_tensor(nm, arr))
        return nm

    def emit(op, extra_inputs):
        counter[0] += 1
        out = f"t{counter[0]}"
        nodes.append(_node(op, [cur[0]] + extra_inputs, [out], name=f"{op}_{counter[0]}"))
        cur[0] = out

    def bn(prefix):
        emit("BatchNormalization", [add_init(f"{prefix}.{n}", state[f"{prefix}.{n}"]) for n in ("weight", "bias", "running_mean", "running_var")])

    def conv(pc, pb, op="Conv"):
        w, b = state[f"{pc}.weight"], state[f"{pc}.bias"]
        if pb is not None and fuse_conv_bn and op == "Conv":
            s = state[f"{pb}.weight"] / np.sqrt(state[f"{pb}.running_var"] + 1e-5)
            w = w * s[:, None, None, None]
            b = (b - state[f"{pb}.running_mean"]) * s + state[f"{pb}.bias"]
            emit(op, [add_init(f"{pc}.weight", w), add_init(f"{pc}.bias", b)])
        else:
            emit(op, [add_init(f"{pc}.weight", w), add_init(f"{pc}.bias", b)])
            if pb is not None:
                bn(pb)
        if pb is not None:
            emit("Relu", [])

    def block(p):
        for i in range(l):
            conv(f"{p}.tfc.H.{i}.0", f"{p}.tfc.H.{i}.1")
        res = cur[0]
        for a, b_ in (("tdf.0", "tdf.1"), ("tdf.3", "tdf.4")):
            emit("MatMul", [add_init(f"{p}.{a}.weight.T", state[f"{p}.{a}.weight"].T)])
            bn(f"{p}.{b_}")
            emit("Relu", [])
        emit("Add", [res])

    n = num_blocks // 2
    conv("first_conv.0", "first_conv.1")
    emit("Transpose", [])
    skips = []
    for i in range(n):
        block(f"encoding_blocks.{i}")
        skips.append(cur[0])
        conv(f"ds.{i}.0", f"ds.{i}.1")
    block("bottleneck_block")
    for i in range(n):
        conv(f"us.{i}.0", f"us.{i}.1", op="ConvTranspose")
        emit("Mul", [skips[-i - 1]])
        block(f"decoding_blocks.{i}")
    emit("Transpose", [])
    conv("final_conv.0", None)
    graph = b"".join(_ld(1, nd) for nd in nodes) + _ld(2, b"convtdfnet") + b"".join(_ld(5, t) for t in inits)
    model = _vi(1, 8) + _ld(2, b"b200sep-tests") + _ld(7, graph)
    with open(path, "wb") as f:
        f.write(model)
