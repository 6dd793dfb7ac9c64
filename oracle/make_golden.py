"""Pin the oracle against the UNMODIFIED reference and (re)generate tests/golden/*.npz.

Runs only in the build container (needs /root/reference; see oracle/ref_shim.py).  For every stage of the
MDX hot path it executes the reference's own code and this repo's restatement (oracle/mdx_oracle.py) on the
same seeded inputs, asserts agreement, and stores the REFERENCE outputs as golden vectors:

  stage                      reference symbol (file:line)
  STFT forward / inverse     uvr_lib_v5/stft.py:20-56, :99-126
  network forward            uvr_lib_v5/mdxnet.py:99-120 (ConvTDFNet, torch CPU fp32)
  run_model / demix          architectures/mdx_separator.py:414-450, :293-412  (MDXSeparator via object.__new__)
  separate() array glue      architectures/mdx_separator.py:152-182, uvr_lib_v5/spec_utils.py:99-115

Usage:  python oracle/make_golden.py [--full]     (--full adds the Inst_HQ_3-sized single-chunk vector, ~20 s)
"""
import argparse
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mdx_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

SMALL = dict(n_fft=1536, hop_length=256, dim_f=768, dim_t=32, segment_size=32, g=8, overlap=0.25, compensate=1.022)
FULL = dict()  # MDXConfig defaults = UVR-MDX-NET-Inst_HQ_3


def ref_net(cfg, weights):
    mdxnet = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.mdxnet")
    net = mdxnet.ConvTDFNet("x", 1e-3, "rmsprop", cfg.dim_c, cfg.dim_f, cfg.dim_t, cfg.n_fft, cfg.hop_length, cfg.num_blocks, cfg.l, cfg.g, cfg.k, cfg.bn, False, 0)
    sd = net.state_dict()
    new = {}
    for k, v in sd.items():
        if k in weights:
            assert tuple(v.shape) == weights[k].shape, (k, v.shape, weights[k].shape)
            new[k] = torch.from_numpy(weights[k])
        else:
            assert k in ("window", "freq_pad") or k.endswith("num_batches_tracked"), k
            new[k] = v
    missing = set(weights) - set(sd)
    assert not missing, missing
    net.load_state_dict(new)
    return net.eval()


def ref_separator(cfg, model_run):
    """A reference MDXSeparator without __init__ (no model files / onnxruntime here): attribute injection."""
    mdx = ref_shim.ref_module("audio_separator.separator.architectures.mdx_separator")
    sep = object.__new__(mdx.MDXSeparator)
    log = logging.getLogger("ref")
    log.setLevel(logging.WARNING)
    sep.logger = log
    sep.log_level = logging.WARNING
    sep.torch_device = torch.device("cpu")
    sep.segment_size = cfg.segment_size
    sep.overlap = cfg.overlap
    sep.batch_size = 1
    sep.hop_length = cfg.hop_length
    sep.enable_denoise = cfg.enable_denoise
    sep.compensate = cfg.compensate
    sep.dim_f, sep.dim_t, sep.n_fft = cfg.dim_f, cfg.dim_t, cfg.n_fft
    sep.model_run = model_run
    return sep


def check(name, a, b, tol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    m = np.isfinite(a) & np.isfinite(b)
    assert (np.isfinite(a) == np.isfinite(b)).all(), name
    err = np.abs(a[m] - b[m]).max()
    scale = np.abs(a[m]).max()
    print(f"  pin {name:34s} max|ref-oracle| = {err:.3e} (ref max {scale:.3e}) tol {tol:.1e}")
    assert err <= tol, (name, err, tol)
    return err


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    stft_mod = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.stft")
    spec_utils_ok = True

    # ---------------- small configuration: every stage, whole demix ----------------
    cfg = O.MDXConfig(**SMALL)
    N = 23000
    mix = O.synth_music(N, seed=1234)
    w = O.make_convtdfnet_weights(cfg, seed=7, out_gain=1.0)
    net = ref_net(cfg, w)

    def model_run_ref(spek):
        with torch.no_grad():
            return net(torch.as_tensor(spek)).numpy()

    def model_run_orc(spek):
        return O.convtdfnet_forward(w, cfg, spek)

    ref_stft = stft_mod.STFT(logging.getLogger("ref"), cfg.n_fft, cfg.hop_length, cfg.dim_f, "cpu")
    chunk = np.zeros((1, 2, cfg.chunk_size), dtype=np.float32)
    chunk[0] = mix[:, : cfg.chunk_size]
    spec_ref = ref_stft(torch.from_numpy(chunk)).numpy()
    spec_orc = O.stft_forward(chunk, cfg.n_fft, cfg.hop_length, cfg.dim_f)
    check("small stft_forward", spec_ref, spec_orc, 2e-4 * max(1.0, np.abs(spec_ref).max()) * 1e-1)
    wav_ref = ref_stft.inverse(torch.from_numpy(spec_ref)).numpy()
    wav_orc = O.stft_inverse(spec_ref, cfg.n_fft, cfg.hop_length)
    check("small stft_inverse", wav_ref, wav_orc, 5e-6)
    net_ref = model_run_ref(spec_ref)
    net_orc = model_run_orc(spec_ref)
    check("small convtdfnet_forward", net_ref, net_orc, 1e-5 * max(1.0, np.abs(net_ref).max()))

    # calibrate the synthetic net's output gain so separated audio is O(1) like a real model's
    gain = 0.5 / max(1e-9, np.abs(ref_stft.inverse(torch.from_numpy(net_ref)).numpy()).max())
    w = O.make_convtdfnet_weights(cfg, seed=7, out_gain=gain)
    net = ref_net(cfg, w)
    net_ref = model_run_ref(spec_ref)

    sep = ref_separator(cfg, model_run_ref)
    mixn = O.normalize(mix, 0.9, 0.0)
    dem_ref = sep.demix(mixn.copy())
    dem_orc = O.demix(mixn.copy(), cfg, model_run_orc)
    check("small demix", dem_ref, dem_orc, 2e-5)
    mm_ref = sep.demix(mixn.copy(), is_match_mix=True)
    mm_orc = O.demix(mixn.copy(), cfg, model_run_orc, is_match_mix=True)
    check("small demix(is_match_mix)", mm_ref, mm_orc, 2e-5)
    # separate() glue restated from mdx_separator.py:152-182 using the reference's demix + normalize
    su = None
    try:
        su = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.spec_utils")
    except Exception as e:  # librosa-dependent module; normalize is 10 lines of numpy
        print("  (spec_utils not importable here: %s; using restated normalize)" % type(e).__name__)
    norm = su.normalize if su is not None else (lambda wave, max_peak, min_peak: O.normalize(wave, max_peak, min_peak))
    peak = np.abs(mix).max()
    mixr = norm(wave=mix.copy(), max_peak=0.9, min_peak=0.0)
    prim_ref = (sep.demix(mixr) * peak).T
    sec_ref = (-prim_ref * cfg.compensate) + mixr.T
    prim_orc, sec_orc = O.separate_arrays(mix, cfg, model_run_orc)
    check("small separate primary", prim_ref, prim_orc, 2e-5)
    check("small separate secondary", sec_ref, sec_orc, 2e-5)
    # denoise branch
    cfg_d = O.MDXConfig(**{**SMALL, "enable_denoise": True})
    sep_d = ref_separator(cfg_d, lambda s: model_run_ref(s.numpy() if hasattr(s, "numpy") else s))
    sep_d.initialize_model_settings()
    den_ref = sep_d.run_model(torch.from_numpy(chunk))
    den_orc = O.run_model(chunk, cfg_d, model_run_orc)
    check("small run_model(denoise)", den_ref, den_orc, 2e-5)

    np.savez_compressed(
        os.path.join(GOLD, "mdx_small.npz"),
        cfg=np.array([cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.dim_t, cfg.segment_size, cfg.g]),
        weights_seed=7,
        out_gain=np.float64(gain),
        mix_seed=1234,
        n_samples=N,
        spec_ref=spec_ref.astype(np.float32),
        istft_ref=wav_ref.astype(np.float32),
        net_ref=net_ref.astype(np.float32),
        demix_ref=dem_ref.astype(np.float32),
        matchmix_ref=mm_ref.astype(np.float32),
        primary_ref=prim_ref.astype(np.float32),
        secondary_ref=sec_ref.astype(np.float32),
        denoise_ref=den_ref.astype(np.float32),
    )
    print("wrote tests/golden/mdx_small.npz")

    # ---------------- ragged / edge-case chunk grids (reference demix with a trivial network) ----------------
    edge = {}
    ident = lambda s: (s.numpy() if hasattr(s, "numpy") else s) * 0.5  # noqa: E731
    for n_edge in (1, 777, cfg.gen_size - 1, cfg.gen_size, cfg.gen_size + 1, 2 * cfg.gen_size + 5):
        m = O.synth_music(max(n_edge, 64), seed=99)[:, :n_edge]
        sep_e = ref_separator(cfg, ident)
        r = sep_e.demix(m.copy())
        o = O.demix(m.copy(), cfg, ident)
        check(f"edge demix N={n_edge}", r, o, 2e-5)
        edge[f"n{n_edge}"] = r.astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "mdx_edge.npz"), **edge)
    print("wrote tests/golden/mdx_edge.npz")

    # ---------------- Inst_HQ_3-sized single chunk (subsampled) ----------------
    if args.full:
        cfgF = O.MDXConfig(**FULL)
        mixF = O.synth_music(cfgF.chunk_size, seed=1234)
        mixF = O.normalize(mixF, 0.9, 0.0)
        chunkF = mixF[None]
        wF = O.make_convtdfnet_weights(cfgF, seed=11, out_gain=1.0)
        netF = ref_net(cfgF, wF)
        stF = stft_mod.STFT(logging.getLogger("ref"), cfgF.n_fft, cfgF.hop_length, cfgF.dim_f, "cpu")
        specF = stF(torch.from_numpy(chunkF)).numpy()
        specF[:, :, :3, :] *= 0
        with torch.no_grad():
            outF = netF(torch.from_numpy(specF)).numpy()
        gainF = 0.5 / np.abs(stF.inverse(torch.from_numpy(outF)).numpy()).max()
        outF = outF * np.float32(gainF)  # the final 1x1 conv is linear in its weight; bias is scaled too, see below
        # rebuild exactly: scale weight only (bias unscaled) -> recompute through the reference for exactness
        wF = O.make_convtdfnet_weights(cfgF, seed=11, out_gain=gainF)
        netF = ref_net(cfgF, wF)
        with torch.no_grad():
            outF = netF(torch.from_numpy(specF)).numpy()
        wavF = stF.inverse(torch.from_numpy(outF)).numpy()
        check("full stft_forward", stF(torch.from_numpy(chunkF)).numpy(), O.stft_forward(chunkF, cfgF.n_fft, cfgF.hop_length, cfgF.dim_f), 5e-3)
        check("full convtdfnet_forward", outF, O.convtdfnet_forward(wF, cfgF, specF), 1e-5 * max(1.0, np.abs(outF).max()))
        check("full stft_inverse", wavF, O.stft_inverse(outF, cfgF.n_fft, cfgF.hop_length), 5e-6)
        np.savez_compressed(
            os.path.join(GOLD, "mdx_full_chunk.npz"),
            weights_seed=11,
            out_gain=np.float64(gainF),
            mix_seed=1234,
            stride=16,
            spec_ref_sub=specF[0, :, ::16, ::4].astype(np.float32),
            net_ref_sub=outF[0, :, ::16, ::4].astype(np.float32),
            wav_ref_sub=wavF[0, :, ::16].astype(np.float32),
            spec_abs_sum=np.float64(np.abs(specF.astype(np.float64)).sum()),
            net_abs_sum=np.float64(np.abs(outF.astype(np.float64)).sum()),
            wav_abs_sum=np.float64(np.abs(wavF.astype(np.float64)).sum()),
        )
        print("wrote tests/golden/mdx_full_chunk.npz")
    print("oracle pinned against the reference: OK")


if __name__ == "__main__":
    main()
