"""GPU parity of the MDXC (MDX23C TFC_TDF_net) path against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden_mdxc.py) and against the oracle.  Tolerance: 1e-4 max-abs on audio (BASELINE gate); network-domain
outputs relative to their max (split-bf16 arithmetic, ~40 sequential contractions)."""
import os

import numpy as np
import pytest
import torch

import mdx_oracle as M
import mdxc_oracle as X

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(lib_built, golden_dir):
    assert torch.cuda.is_available()
    from audio_separator.separator.b200 import engine

    z = np.load(os.path.join(golden_dir, "mdxc_small.npz"))
    n_fft, hop, dim_f, dim_t, n_scales, c, g, bn, overlap = (int(v) for v in z["cfg"])
    cfg = X.MDXCConfig(n_fft=n_fft, hop_length=hop, dim_f=dim_f, dim_t=dim_t, num_scales=n_scales, num_channels_model=c, growth=g, bottleneck_factor=bn, overlap=overlap)
    w = X.make_weights(cfg, seed=int(z["weights_seed"]), out_gain=float(z["out_gain"]))
    net = engine.TfcNet(w, cfg.dim_f, cfg.dim_t, cfg.num_subbands, 2, cfg.num_scales, cfg.num_blocks_per_scale, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor, cfg.num_targets, max_batch=2)
    eng = engine.MdxcEngine(net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.dim_t, cfg.overlap)
    return z, cfg, w, net, eng


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def test_tfcnet_spec_vs_reference_golden(setup):
    z, cfg, w, net, eng = setup
    spec = z["spec_in"]  # (1,4,F,T) reference layout
    got = net.forward_spec(dev(spec.transpose(0, 1, 3, 2))).cpu().numpy()  # (1,S,4,T,F)
    got = got.transpose(0, 1, 2, 4, 3).reshape(1, cfg.num_targets * 4, cfg.dim_f, cfg.dim_t)
    ref = z["spec_out"]
    err = np.abs(got.astype(np.float64) - ref).max()
    assert err <= 3e-4 * np.abs(ref).max(), (err, np.abs(ref).max())
    # batch of 3 through max_batch 2
    rng = np.random.default_rng(0)
    x3 = (rng.standard_normal((3, 4, cfg.dim_f, cfg.dim_t)) * 2).astype(np.float32)
    ref3 = X.net_forward_spec(w, cfg, x3, dtype="float64")
    got3 = net.forward_spec(dev(x3.transpose(0, 1, 3, 2))).cpu().numpy().transpose(0, 1, 2, 4, 3).reshape(ref3.shape)
    assert np.abs(got3 - ref3).max() <= 3e-4 * np.abs(ref3).max()


def test_model_run_and_demix_vs_reference_golden(setup):
    z, cfg, w, net, eng = setup
    mix = M.normalize(M.synth_music(int(z["n_samples"]), seed=int(z["mix_seed"])), 0.9, 0.0)
    out = eng.model_run(dev(mix[None, :, : cfg.chunk_size])).cpu().numpy()
    assert out.shape == z["chunk_out"].shape == (1, 2, 2, cfg.chunk_size)
    assert np.abs(out - z["chunk_out"]).max() <= 1e-4
    dem = eng.demix_device(dev(mix)).cpu().numpy()
    assert dem.shape == z["demix_ref"].shape == (2, 2, mix.shape[1])  # stems x channels x samples, sample count exact
    assert np.abs(dem - z["demix_ref"]).max() <= 1e-4


@pytest.mark.parametrize("n", [1, 959, 3840, 3841])
def test_demix_edge_lengths_vs_oracle(setup, n):
    z, cfg, w, net, eng = setup
    m = M.synth_music(max(64, n), seed=5)[:, :n]
    assert eng.grid(n) == X.chunk_grid(n, cfg)
    ref = X.demix(m, cfg, lambda x: X.net_forward(w, cfg, x))
    got = eng.demix_device(dev(m)).cpu().numpy()
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4


def test_mdxc_plugin_end_to_end(setup, tmp_path):
    import json
    import wave

    from audio_separator.separator import Separator

    z, cfg, w, net, eng = setup
    np.savez(tmp_path / "tiny-mdx23c.npz", **w)
    yaml_cfg = {
        "audio": {"n_fft": cfg.n_fft, "hop_length": cfg.hop_length, "dim_f": cfg.dim_f, "num_channels": 2, "chunk_size": cfg.chunk_size, "sample_rate": 44100},
        "model": {"norm": "InstanceNorm", "act": "gelu", "num_subbands": 4, "num_scales": cfg.num_scales, "scale": [2, 2], "num_blocks_per_scale": 2,
                  "num_channels": cfg.num_channels_model, "growth": cfg.growth, "bottleneck_factor": cfg.bottleneck_factor},
        "training": {"instruments": ["Vocals", "Instrumental"], "target_instrument": None},
        "inference": {"dim_t": cfg.dim_t},
    }
    (tmp_path / "tiny-mdx23c.json").write_text(json.dumps(yaml_cfg))
    mix = M.synth_music(12000, seed=9)
    pcm = (mix.T * 32767).astype("<i2")
    with wave.open(str(tmp_path / "song.wav"), "wb") as wf:
        wf.setnchannels(2); wf.setsampwidth(2); wf.setframerate(44100); wf.writeframes(pcm.tobytes())
    sep = Separator(model_file_dir=str(tmp_path), output_dir=str(tmp_path / "out"), mdxc_params={"overlap": cfg.overlap, "batch_size": 2, "segment_size": cfg.dim_t})
    sep.load_model("tiny-mdx23c.npz")
    assert not sep.model_instance.override_model_segment_size
    files = sep.separate(str(tmp_path / "song.wav"))
    # a clip shorter than 10 s switches the plugin to dim_t = segment_size, like the reference (mdxc_separator.py:137-143); here both are 16
    assert sep.model_instance.override_model_segment_size and sep.model_instance.engine.dim_t == cfg.dim_t
    assert files == ["song_(Instrumental)_tiny-mdx23c.wav", "song_(Vocals)_tiny-mdx23c.wav"]  # two-stem dict: the secondary file first (mdxc_separator.py:186-214)
    loaded = M.normalize(pcm.astype(np.float32).T / 32768.0, 0.9, 0.0)
    ref = X.demix(loaded, cfg, lambda x: X.net_forward(w, cfg, x))  # rows in training.instruments order: Vocals, Instrumental
    for fname, stem in zip(files, (ref[1], ref[0])):
        with wave.open(str(tmp_path / "out" / fname)) as wf:
            assert wf.getnframes() == 12000
            got = np.frombuffer(wf.readframes(12000), dtype="<i2").astype(np.int32)
        want = M.to_pcm16(stem.T.copy(), 0.9, 0.0).astype(np.int32)
        assert np.abs(got - want).max() <= 4
