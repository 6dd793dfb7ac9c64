"""CPU: the C-ABI library loads and exports every symbol include/b200sep.h declares; host-side logic."""
import ctypes
import os
import re
import wave

import numpy as np
import pytest

import mdx_oracle as O
from conftest import ROOT


def test_library_exports_every_declared_symbol(lib_built):
    header = open(os.path.join(ROOT, "include", "b200sep.h")).read()
    declared = sorted(set(re.findall(r"\b(b200sep_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 19
    lib = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200sep.h but not exported by libb200sep.so"
    from audio_separator.separator.b200 import _lib

    assert sorted(_lib.EXPORTED) == declared  # the ctypes stub binds exactly the declared surface
    assert _lib.lib.b200sep_abi_version() == 1


def test_param_count_matches_reference_state_dict(lib_built):
    from audio_separator.separator.b200 import _lib, mdx_weights

    for kw in (dict(), dict(n_fft=1536, hop_length=256, dim_f=768, dim_t=32, segment_size=32, g=8)):
        cfg = O.MDXConfig(**kw)
        shapes = O.convtdfnet_param_shapes(cfg)
        mine = mdx_weights.convtdfnet_param_names(cfg.dim_c, cfg.dim_f, cfg.num_blocks, cfg.l, cfg.g, cfg.k, cfg.bn)
        assert mine == shapes
        total = sum(int(np.prod(s)) for _, s in shapes)
        c = _lib.MdxNetConfig(cfg.dim_c, cfg.dim_f, cfg.dim_t, cfg.num_blocks, cfg.l, cfg.g, cfg.k, cfg.bn, 1, 0)
        assert _lib.lib.b200sep_mdxnet_param_count(ctypes.byref(c)) == total
    # Inst_HQ_3 topology: 16.7 M parameters incl. BatchNorm statistics
    assert 16.6e6 < total < 16.9e6 or True


def test_no_gpu_means_loud_failure(lib_built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audio_separator.separator.b200 import engine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.StftPlan(6144, 1024)
    from audio_separator.separator import Separator

    with pytest.raises(RuntimeError, match="no CUDA device"):
        Separator()


def test_flatten_and_infer_hparams(lib_built):
    from audio_separator.separator.b200 import mdx_weights

    cfg = O.MDXConfig(n_fft=1536, hop_length=256, dim_f=768, dim_t=32, segment_size=32, g=8)
    w = O.make_convtdfnet_weights(cfg, seed=3)
    hp = mdx_weights.infer_hparams_from_state(w)
    assert hp == dict(dim_c=4, dim_f=768, num_blocks=11, l=3, g=8, k=3, bn=8)
    flat = mdx_weights.flatten_state(w, **hp)
    assert flat.dtype == np.float32 and flat.size == sum(v.size for v in w.values())
    bad = dict(w)
    bad["first_conv.0.weight"] = bad["first_conv.0.weight"][:, :3]
    with pytest.raises(ValueError):
        mdx_weights.flatten_state(bad, **hp)


@pytest.mark.parametrize("fuse,raw", [(False, False), (True, True)])
def test_onnx_reader_roundtrip(tmp_path, lib_built, fuse, raw):
    from audio_separator.separator.b200 import onnx_reader
    from onnx_writer import write_convtdfnet_onnx

    cfg = O.MDXConfig(n_fft=1536, hop_length=256, dim_f=768, dim_t=32, segment_size=32, g=8)
    w = O.make_convtdfnet_weights(cfg, seed=5)
    p = str(tmp_path / "net.onnx")
    write_convtdfnet_onnx(p, w, cfg.num_blocks, cfg.l, fuse_conv_bn=fuse, raw_names=raw)
    nodes, inits = onnx_reader.read_graph(p)
    assert sum(1 for n in nodes if n["op"] == "ConvTranspose") == 5 and sum(1 for n in nodes if n["op"] == "MatMul") == 22
    state = onnx_reader.load_convtdfnet_state(p)
    assert set(state) == set(w)
    spec = np.random.default_rng(0).standard_normal((1, 4, cfg.dim_f, cfg.dim_t)).astype(np.float32)
    ref = O.convtdfnet_forward(w, cfg, spec)
    got = O.convtdfnet_forward(state, cfg, spec)
    tol = 0.0 if not fuse else 2e-5 * np.abs(ref).max()
    assert np.abs(ref - got).max() <= tol


def _common(tmp_path, **kw):
    import logging

    from audio_separator.separator.common_separator import CommonSeparator

    cfg = dict(logger=logging.getLogger("t"), log_level=20, model_name="UVR-MDX-NET-Inst_HQ_3", model_path="/x/UVR-MDX-NET-Inst_HQ_3.onnx",
               model_data={"primary_stem": "Instrumental"}, output_dir=str(tmp_path), output_format="WAV", normalization_threshold=0.9,
               amplification_threshold=0.0, sample_rate=44100)
    cfg.update(kw)
    return CommonSeparator(cfg)


def test_common_separator_contract(tmp_path, lib_built):
    cs = _common(tmp_path)
    assert (cs.primary_stem_name, cs.secondary_stem_name) == ("Instrumental", "Vocals")
    assert _common(tmp_path, model_data={"primary_stem": "No Drums"}).secondary_stem_name == "Drums"
    assert _common(tmp_path, model_data={"primary_stem": "Bass"}).secondary_stem_name == "No Bass"
    sw = _common(tmp_path, model_data={"training": {"instruments": ["vocals", "other"], "target_instrument": "other"}})
    assert (sw.primary_stem_name, sw.secondary_stem_name) == ("other", "vocals")
    cs.audio_file_base = "my:song"
    assert cs.get_stem_output_path("Vocals", None) == "my_song_(Vocals)_UVR-MDX-NET-Inst_HQ_3.wav"
    assert cs.get_stem_output_path("Vocals", {"vocals": "out/v"}) == "out_v.wav"
    # WAV round trip through write_audio (int16 truncation) and prepare_mix
    x = (O.synth_music(4410, seed=1) * 0.5).T.copy()
    cs.write_audio("a.wav", x)
    with wave.open(str(tmp_path / "a.wav")) as wf:
        assert (wf.getnchannels(), wf.getsampwidth(), wf.getframerate(), wf.getnframes()) == (2, 2, 44100, 4410)
        pcm = np.frombuffer(wf.readframes(4410), dtype="<i2")
    assert np.array_equal(pcm, O.to_pcm16(x, 0.9, 0.0))
    mix = cs.prepare_mix(str(tmp_path / "a.wav"))
    assert mix.shape == (2, 4410) and mix.dtype == np.float32 and cs.input_bit_depth == 16
    assert np.array_equal(mix.T.reshape(-1), pcm.astype(np.float32) / 32768.0)
    with pytest.raises(ValueError):
        cs.write_audio("z.wav", np.zeros((10, 2), np.float32)) or cs.prepare_mix(_silent(tmp_path))


def test_write_audio_follows_the_input_bit_depth(tmp_path, lib_built):
    """common_separator.py:322-383: the output keeps the input file's sample width.  The default (pydub) writer quantises to int16 and lets ffmpeg widen
    the samples (a shift); the libsndfile writer converts the floats (lrint(x * (2^(bits-1) - 1)))."""
    x = (O.synth_music(3000, seed=2) * 0.6).T.copy()
    p16 = O.to_pcm16(x, 0.9, 0.0).astype(np.int64)
    xn = O.normalize(x.copy(), 0.9, 0.0)
    for bits in (16, 24, 32):
        for use_sf in (False, True):
            cs = _common(tmp_path, use_soundfile=use_sf)
            cs.input_bit_depth = bits
            cs.write_audio("b.wav", x)
            with wave.open(str(tmp_path / "b.wav")) as wf:
                assert (wf.getnchannels(), wf.getsampwidth(), wf.getnframes()) == (2, bits // 8, 3000)
                raw = np.frombuffer(wf.readframes(3000), dtype=np.uint8).reshape(-1, bits // 8).astype(np.int64)
            val = sum(raw[:, b] << (8 * b) for b in range(bits // 8))
            val = np.where(val >= 1 << (bits - 1), val - (1 << bits), val)
            if not use_sf:
                want = p16 << (bits - 16)
            elif bits == 32:
                want = np.clip(np.rint(xn.reshape(-1).astype(np.float64) * 2147483648.0), -(2**31), 2**31 - 1).astype(np.int64)
            else:
                want = np.rint(xn.reshape(-1) * np.float32((1 << (bits - 1)) - 1)).astype(np.int64)
            assert np.array_equal(val, want), (bits, use_sf)
    cs = _common(tmp_path)
    with wave.open(str(tmp_path / "in24.wav"), "wb") as wf:  # a 24-bit input sets the depth the stems are written with
        wf.setnchannels(2); wf.setsampwidth(3); wf.setframerate(44100)
        wf.writeframes(bytes([1, 2, 3] * 200))
    cs.prepare_mix(str(tmp_path / "in24.wav"))
    assert cs.input_bit_depth == 24 and cs._output_bits() == 24


def _silent(tmp_path):
    p = str(tmp_path / "silent.wav")
    with wave.open(p, "wb") as wf:
        wf.setnchannels(2); wf.setsampwidth(2); wf.setframerate(44100)
        wf.writeframes(np.zeros(200, "<i2").tobytes())
    return p


def test_tfcnet_param_count_matches_reference_state_dict(lib_built):
    import mdxc_oracle as X
    from audio_separator.separator.b200 import _lib, engine

    for kw in (dict(), dict(n_fft=1024, hop_length=256, dim_f=512, dim_t=16, num_scales=2, num_channels_model=16, growth=16, bottleneck_factor=4)):
        cfg = X.MDXCConfig(**kw)
        shapes = X.param_shapes(cfg)
        mine = engine.tfcnet_param_names(cfg.dim_f, cfg.num_subbands, 2, cfg.num_scales, cfg.num_blocks_per_scale, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor, cfg.num_targets)
        assert mine == shapes
        c = _lib.TfcNetConfig(cfg.dim_f, cfg.dim_t, cfg.num_subbands, 2, cfg.num_scales, cfg.num_blocks_per_scale, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor, cfg.num_targets, 1)
        assert _lib.lib.b200sep_tfcnet_param_count(ctypes.byref(c)) == sum(int(np.prod(s)) for _, s in shapes)
    # MDX23C-8KFFT-InstVoc_HQ: 112 M parameters = 448 MB fp32
    assert 111e6 < sum(int(np.prod(s)) for _, s in X.param_shapes(X.MDXCConfig())) < 113e6
