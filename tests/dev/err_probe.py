"""Development probe: full-size single-chunk error of the two precision modes vs the committed reference golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("python-audio-separator_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import mdx_oracle as O
from audio_separator.separator.b200 import engine, mdx_weights
z = np.load(os.path.join(ROOT, "tests/golden/mdx_full_chunk.npz"))
cfg = O.MDXConfig()
w = O.make_convtdfnet_weights(cfg, seed=int(z["weights_seed"]), out_gain=float(z["out_gain"]))
mix = O.normalize(O.synth_music(cfg.chunk_size, seed=int(z["mix_seed"])), 0.9, 0.0)
hp = mdx_weights.infer_hparams_from_state(w)
flat = mdx_weights.flatten_state(w, **hp)
for precision in (0, 1):
    net = engine.MdxNet(flat, dim_t=cfg.dim_t, max_batch=1, precision=precision, **hp)
    e = engine.MdxEngine(net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.segment_size, cfg.overlap, cfg.compensate)
    spec = e.plan.forward(torch.as_tensor(mix[None]).cuda(), cfg.dim_f, 3, engine.LAYOUT_CFT)
    out = net.forward(spec).cpu().numpy()[0, :, ::16, ::4]
    ref = z["net_ref_sub"]
    d = np.abs(out.astype(np.float64) - ref)
    wav = e.run_model(torch.as_tensor(mix[None]).cuda()).cpu().numpy()[0, :, ::16]
    dw = np.abs(wav.astype(np.float64) - z["wav_ref_sub"])
    print(f"precision={precision}: net max|err|={d.max():.3e} (ref max {np.abs(ref).max():.2f}, rms {ref.std():.3f}) rms err={np.sqrt((d**2).mean()):.3e} | wav max|err|={dw.max():.3e} rms={np.sqrt((dw**2).mean()):.3e} (wav peak {np.abs(z['wav_ref_sub']).max():.3f})")
    del net, e
