"""Development probe: full-size MDX23C-8KFFT topology (112 M params): one chunk vs the CPU oracle + forward timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("python-audio-separator_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import mdx_oracle as M, mdxc_oracle as X
from audio_separator.separator.b200 import engine
torch.set_num_threads(32)
cfg = X.MDXCConfig()
w = X.make_weights(cfg, seed=4, out_gain=1.0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = engine.TfcNet(w, cfg.dim_f, cfg.dim_t, cfg.num_subbands, 2, cfg.num_scales, cfg.num_blocks_per_scale, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor, cfg.num_targets, max_batch=B)
print(f"device bytes: {net.device_bytes/2**30:.2f} GiB (batch {B})")
eng = engine.MdxcEngine(net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.dim_t, cfg.overlap)
mix = M.normalize(M.synth_music(cfg.chunk_size, seed=8), 0.9, 0.0)
x = torch.as_tensor(np.tile(mix[None], (B, 1, 1))).cuda()
y = eng.model_run(x); torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(3): y = eng.model_run(x)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 3
hop_s = cfg.hop_size / 44100
print(f"model_run batch {B}: {ms:.2f} ms  -> {ms/B:.2f} ms/chunk; RTF at overlap {cfg.overlap}: {hop_s / (ms / B / 1e3):.1f}; tensor rate {2434.1e9 * B / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic")
if "--check" in sys.argv:
    t0 = time.time(); ref = X.net_forward(w, cfg, mix[None]); print(f"oracle CPU forward {time.time() - t0:.1f} s")
    got = y[:1].cpu().numpy()
    scale = np.abs(ref).max()
    print(f"max|gpu-oracle| = {np.abs(got - ref).max():.3e} (ref max {scale:.3e}) -> relative {np.abs(got - ref).max() / scale:.2e}")
