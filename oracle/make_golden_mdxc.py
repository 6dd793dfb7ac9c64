"""Pin oracle/mdxc_oracle.py against the UNMODIFIED reference TFC_TDF_net + MDXCSeparator.demix and write
tests/golden/mdxc_small.npz (build container only; see oracle/ref_shim.py)."""
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mdx_oracle as M  # noqa: E402
import mdxc_oracle as X  # noqa: E402
import ref_shim  # noqa: E402
from make_golden import check  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
SMALL = dict(n_fft=1024, hop_length=256, dim_f=512, dim_t=16, num_scales=2, num_channels_model=16, growth=16, bottleneck_factor=4, overlap=4)


class AD(dict):
    __getattr__ = dict.__getitem__


def ref_cfg(cfg):
    return AD(
        audio=AD(n_fft=cfg.n_fft, hop_length=cfg.hop_length, dim_f=cfg.dim_f, num_channels=cfg.num_channels, chunk_size=cfg.chunk_size, sample_rate=44100),
        model=AD(norm="InstanceNorm", act="gelu", num_subbands=cfg.num_subbands, num_scales=cfg.num_scales, scale=list(cfg.scale), num_blocks_per_scale=cfg.num_blocks_per_scale,
                 num_channels=cfg.num_channels_model, growth=cfg.growth, bottleneck_factor=cfg.bottleneck_factor),
        training=AD(instruments=list(cfg.instruments), target_instrument=cfg.target_instrument),
        inference=AD(dim_t=cfg.dim_t),
    )


def ref_net(cfg, w):
    mod = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.tfc_tdf_v3")
    net = mod.TFC_TDF_net(ref_cfg(cfg), "cpu")
    sd = net.state_dict()
    assert set(sd) == set(w), (set(sd) ^ set(w))
    assert [k for k in sd] == [n for n, _ in X.param_shapes(cfg)], "state_dict order differs from the oracle's param order"
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return net.eval()


def main():
    cfg = X.MDXCConfig(**SMALL)
    w = X.make_weights(cfg, seed=3)
    net = ref_net(cfg, w)
    N = 9000
    mix = M.normalize(M.synth_music(N, seed=77), 0.9, 0.0)
    chunk = mix[None, :, : cfg.chunk_size]
    with torch.no_grad():
        y_ref = net(torch.from_numpy(chunk)).numpy()
    gain = 0.4 / np.abs(y_ref).max()
    w = X.make_weights(cfg, seed=3, out_gain=gain)
    net = ref_net(cfg, w)
    with torch.no_grad():
        y_ref = net(torch.from_numpy(chunk)).numpy()
    y_orc = X.net_forward(w, cfg, chunk)
    check("mdxc net forward (wave->stems)", y_ref, y_orc, 2e-5)
    spec = M.stft_forward(chunk, cfg.n_fft, cfg.hop_length, cfg.dim_f)
    spec_out = X.net_forward_spec(w, cfg, spec)

    # reference demix: MDXCSeparator without __init__
    mdxc = ref_shim.ref_module("audio_separator.separator.architectures.mdxc_separator")
    sep = object.__new__(mdxc.MDXCSeparator)
    sep.logger = logging.getLogger("ref")
    sep.is_roformer = False
    sep.model_run = net
    sep.override_model_segment_size = False
    sep.model_data_cfgdict = ref_cfg(cfg)
    sep.overlap = cfg.overlap
    sep.batch_size = 2
    sep.torch_device = torch.device("cpu")
    sep.pitch_shift = 0
    sep.is_primary_stem_main_target = False
    sep.primary_stem_name, sep.secondary_stem_name = "Vocals", "Instrumental"
    src = sep.demix(mix.copy())
    dem_ref = np.stack([src[k] for k in cfg.instruments])
    dem_orc = X.demix(mix, cfg, lambda x: X.net_forward(w, cfg, x))
    check("mdxc demix", dem_ref, dem_orc, 2e-5)
    for n_edge in (1, cfg.hop_size - 1, cfg.chunk_size, cfg.chunk_size + 1):
        m = M.synth_music(max(64, n_edge), seed=5)[:, :n_edge]
        r = np.stack([sep.demix(m.copy())[k] for k in cfg.instruments])
        check(f"mdxc demix N={n_edge}", r, X.demix(m, cfg, lambda x: X.net_forward(w, cfg, x)), 2e-5)
    np.savez_compressed(
        os.path.join(GOLD, "mdxc_small.npz"),
        cfg=np.array([cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.dim_t, cfg.num_scales, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor, cfg.overlap]),
        weights_seed=3, out_gain=np.float64(gain), mix_seed=77, n_samples=N,
        spec_in=spec.astype(np.float32), spec_out=spec_out.astype(np.float32), chunk_out=y_ref.astype(np.float32), demix_ref=dem_ref.astype(np.float32),
    )
    print("wrote tests/golden/mdxc_small.npz; oracle pinned: OK")


if __name__ == "__main__":
    main()
