"""Pin oracle/roformer_oracle.py against the UNMODIFIED reference BSRoformer + the Roformer branch of MDXCSeparator.demix and write
tests/golden/roformer_small.npz.  rotary-embedding-torch is not installed: a stand-in module implementing its published algorithm
(the oracle's apply_rotary) is registered under that name before the reference module is imported."""
import logging
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mdx_oracle as M  # noqa: E402
import ref_shim  # noqa: E402
import roformer_oracle as R  # noqa: E402
from make_golden_vr import check  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
SMALL = dict(dim=32, depth=2, time_transformer_depth=1, freq_transformer_depth=2, freqs_per_bands=(2, 2, 4, 4, 8, 12, 16, 17), dim_head=8, heads=4, stft_n_fft=128,
             stft_hop_length=32, stft_win_length=128, dim_t=65, overlap=8)


class _Rotary(nn.Module):  # rotary_embedding_torch.RotaryEmbedding(dim) with its defaults
    def __init__(self, dim, theta=10000):
        super().__init__()
        self.freqs = nn.Parameter(torch.from_numpy(R.rotary_freqs(dim, theta)), requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        assert seq_dim == -2
        return R.apply_rotary(t, self.freqs)


def ref_model(cfg, w):
    if "rotary_embedding_torch" not in sys.modules:
        sys.modules["rotary_embedding_torch"] = types.SimpleNamespace(RotaryEmbedding=_Rotary)
    bs = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.roformer.bs_roformer")
    m = bs.BSRoformer(**cfg.kwargs()).eval()
    sd = m.state_dict()
    names = [n for n, _ in R.param_shapes(cfg)]
    assert list(sd) == names, [(a, b) for a, b in zip(sd, names) if a != b][:5]
    for (n, s), v in zip(R.param_shapes(cfg), sd.values()):
        assert tuple(v.shape) == tuple(s), (n, v.shape, s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return m


def main():
    ref_shim.install()
    out = {}
    cfg = R.BSRoformerConfig(**SMALL)
    w = R.make_weights(cfg, seed=3)
    model = ref_model(cfg, w)
    N = int(2.6 * cfg.chunk_size)
    mix = M.synth_music(N, seed=51)
    chunk = mix[None, :, : cfg.chunk_size]
    with torch.no_grad():
        y_ref = model(torch.from_numpy(chunk)).numpy()
    check("BSRoformer.forward (stereo, 1 stem)", y_ref, R.forward(w, cfg, chunk), 2e-5)
    with torch.no_grad():
        yb_ref = model(torch.from_numpy(np.stack([mix[:, :1500], mix[:, 700:2200]]))).numpy()
    check("BSRoformer.forward (batch 2, short input)", yb_ref, R.forward(w, cfg, np.stack([mix[:, :1500], mix[:, 700:2200]])), 2e-5)
    out.update(forward_ref=y_ref, forward_short_ref=yb_ref)
    cfg2 = R.BSRoformerConfig(**dict(SMALL, num_stems=2, mask_estimator_depth=3, depth=1))
    w2 = R.make_weights(cfg2, seed=4)
    model2 = ref_model(cfg2, w2)
    with torch.no_grad():
        y2_ref = model2(torch.from_numpy(chunk)).numpy()
    check("BSRoformer.forward (2 stems, mask depth 3)", y2_ref, R.forward(w2, cfg2, chunk), 2e-5)
    out["forward_2stem_ref"] = y2_ref
    # ---- MDXCSeparator.demix, Roformer branch
    ms = ref_shim.ref_module("audio_separator.separator.architectures.mdxc_separator")

    class _Cfg(dict):
        __getattr__ = dict.__getitem__

    def run_demix(model_, cfg_, instruments, target, overlap):
        sep = object.__new__(ms.MDXCSeparator)
        sep.logger = logging.getLogger("ref")
        sep.pitch_shift, sep.is_roformer, sep.override_model_segment_size = 0, True, False
        sep.overlap, sep.sample_rate = overlap, 44100
        sep.model_run = model_
        sep.model_data_cfgdict = _Cfg(inference=_Cfg(dim_t=cfg_.dim_t), model=_Cfg(stft_hop_length=cfg_.stft_hop_length), audio=_Cfg(sample_rate=44100, hop_length=cfg_.stft_hop_length),
                                      training=_Cfg(instruments=instruments, target_instrument=target))
        sep.is_primary_stem_main_target = False
        sep.primary_stem_name, sep.secondary_stem_name = (target or instruments[0]), "Rest"
        return sep.demix(mix.copy())

    d_ref = run_demix(model, cfg, ["Vocals", "Instrumental"], "Vocals", 8)
    fn = lambda c: R.forward(w, cfg, c)  # noqa: E731
    d_orc = R.demix(mix, cfg, fn, n_instruments=2)
    check("demix Roformer branch (single target; the result tensor has len(instruments) rows)", np.asarray(d_ref), d_orc[0], 2e-5)
    # an overlap (seconds) small enough to give a step below the chunk size: hamming-weighted overlap-add
    cfg_o = R.BSRoformerConfig(**dict(SMALL, overlap=0.03))
    d_ref_o = run_demix(model, cfg_o, ["Vocals", "Instrumental"], "Vocals", 0.03)
    d_orc_o = R.demix(mix, cfg_o, fn, n_instruments=2)
    check(f"demix Roformer branch (step {cfg_o.step} < chunk {cfg_o.chunk_size})", np.asarray(d_ref_o), d_orc_o[0], 2e-5)
    d2_ref = run_demix(model2, cfg2, ["Vocals", "Instrumental"], None, 8)
    d2_orc = R.demix(mix, cfg2, lambda c: R.forward(w2, cfg2, c), n_instruments=2)
    check("demix Roformer branch (2 stems)", np.stack([d2_ref["Vocals"], d2_ref["Instrumental"]]), d2_orc, 2e-5)
    out.update(mix_seed=51, n_samples=N, demix_ref=np.asarray(d_ref), demix_overlap_ref=np.asarray(d_ref_o), demix_2stem_ref=np.stack([d2_ref["Vocals"], d2_ref["Instrumental"]]))
    # ---- Mel-Band Roformer (librosa.filters.mel is absent: the oracle's restatement of it is injected; only the SUPPORT of the filters matters)
    sys.modules["librosa"].filters = types.SimpleNamespace(mel=lambda sr, n_fft, n_mels: R.mel_filter_bank(sr, n_fft, n_mels))
    sys.modules["librosa.filters"] = sys.modules["librosa"].filters
    mb = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.roformer.mel_band_roformer")
    MSMALL = dict(dim=32, depth=2, time_transformer_depth=1, freq_transformer_depth=1, num_bands=12, dim_head=8, heads=4, mask_estimator_depth=2, stft_n_fft=128, stft_hop_length=32,
                  stft_win_length=128, dim_t=65)
    mcfg = R.MelBandRoformerConfig(**MSMALL)
    mw = R.make_mel_weights(mcfg, seed=6)
    mm = mb.MelBandRoformer(**mcfg.kwargs()).eval()
    msd = mm.state_dict()
    mnames = [n for n, _ in R.mel_param_shapes(mcfg)]
    assert list(msd) == mnames, [(a, b) for a, b in zip(msd, mnames) if a != b][:5]
    for (n, sh), v in zip(R.mel_param_shapes(mcfg), msd.values()):
        assert tuple(v.shape) == tuple(sh), (n, v.shape, sh)
    mm.load_state_dict({k: torch.from_numpy(v) for k, v in mw.items()})
    fpb, fidx, nfpb, nbpf = R.mel_band_layout(mcfg)
    assert torch.equal(mm.freq_indices, torch.from_numpy(fidx)) and torch.equal(mm.num_bands_per_freq, torch.from_numpy(nbpf))
    with torch.no_grad():
        my_ref = mm(torch.from_numpy(chunk)).numpy()
    check("MelBandRoformer.forward (stereo, 1 stem, overlapping bands)", my_ref, R.forward_mel(mw, mcfg, chunk), 2e-5)
    mcfg2 = R.MelBandRoformerConfig(**dict(MSMALL, num_stems=2, mask_estimator_depth=1, depth=1, freq_transformer_depth=2))
    mw2 = R.make_mel_weights(mcfg2, seed=7)
    mm2 = mb.MelBandRoformer(**mcfg2.kwargs()).eval()
    mm2.load_state_dict({k: torch.from_numpy(v) for k, v in mw2.items()})
    with torch.no_grad():
        my2_ref = mm2(torch.from_numpy(chunk)).numpy()
    check("MelBandRoformer.forward (2 stems, mask depth 1)", my2_ref, R.forward_mel(mw2, mcfg2, chunk), 2e-5)
    md_ref = run_demix(mm, mcfg, ["Vocals", "Instrumental"], "Vocals", 8)
    md_orc = R.demix(mix, mcfg, lambda c: R.forward_mel(mw, mcfg, c), n_instruments=2)
    check("demix Roformer branch with a Mel-Band model", np.asarray(md_ref), md_orc[0], 2e-5)
    out.update(mel_forward_ref=my_ref, mel_forward_2stem_ref=my2_ref, mel_demix_ref=np.asarray(md_ref))
    np.savez_compressed(os.path.join(GOLD, "roformer_small.npz"), **{k: (np.asarray(v, dtype=np.float32) if isinstance(v, np.ndarray) else v) for k, v in out.items()})
    print("wrote tests/golden/roformer_small.npz; oracle pinned: OK (rotary-embedding-torch restated, see header)")


if __name__ == "__main__":
    main()
