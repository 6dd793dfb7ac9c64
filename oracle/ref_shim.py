"""Import shim for the UNMODIFIED reference modules under /root/reference (build container only).

TEST INFRASTRUCTURE -- never imported by the product path.  The reference package cannot be imported
normally here (librosa / onnxruntime / soundfile / pydub ... are absent), so this
registers namespace stubs for the package __init__ files and empty stand-ins for the missing third-party
wheels, after which the reference's own hot-path modules (uvr_lib_v5/stft.py, architectures/mdx_separator.py,
uvr_lib_v5/mdxnet.py, tfc_tdf_v3.py, demucs/*, vr_network/*) import and execute unmodified on CPU.

Used by oracle/make_golden.py to (a) pin the oracle restatement and (b) generate tests/golden/*.npz.
/root/reference does not exist on the GPU box; nothing at test/bench run time imports this file.
"""
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("B200SEP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "audio_separator", "separator"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _ns(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def install():
    """Idempotently install the stubs; returns the reference root."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    if "audio_separator" in sys.modules and getattr(sys.modules["audio_separator"], "_b200_shim", False):
        return REF_ROOT
    import torch.nn as nn

    base = os.path.join(REF_ROOT, "audio_separator")
    pkg = _ns("audio_separator", base)
    pkg._b200_shim = True
    _ns("audio_separator.separator", os.path.join(base, "separator"))
    _ns("audio_separator.separator.uvr_lib_v5", os.path.join(base, "separator", "uvr_lib_v5"))
    _ns("audio_separator.separator.architectures", os.path.join(base, "separator", "architectures"))

    class _Any:
        def __init__(self, *a, **k):
            pass

    if "ml_collections" not in sys.modules:
        class _ConfigDict(dict):
            __getattr__ = dict.__getitem__

        _stub("ml_collections", ConfigDict=_ConfigDict)
    for name in ("librosa", "soundfile", "audioread", "onnx", "onnxruntime", "onnx2torch", "julius", "samplerate", "resampy"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "pydub" not in sys.modules:
        _stub("pydub", AudioSegment=_Any)
    if "diffq" not in sys.modules:
        _stub("diffq", DiffQuantizer=_Any, UniformQuantizer=_Any, restore_quantized_state=lambda *a, **k: None)
    if "pytorch_lightning" not in sys.modules:
        _stub("pytorch_lightning", LightningModule=nn.Module)
    if "tqdm" not in sys.modules:
        _stub("tqdm", tqdm=lambda x, *a, **k: x)
    return REF_ROOT


def ref_module(dotted):
    """import_module for a reference module, e.g. 'audio_separator.separator.uvr_lib_v5.stft'."""
    install()
    return importlib.import_module(dotted)
