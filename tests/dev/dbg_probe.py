"""Development probe: per-category device time of one full-size forward (batch 4, or argv[1]) under B200SEP_DBG switches."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("python-audio-separator_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import mdx_oracle as O
from audio_separator.separator.b200 import engine, mdx_weights
cfg = O.MDXConfig()
w = O.make_convtdfnet_weights(cfg, seed=11)
hp = mdx_weights.infer_hparams_from_state(w)
BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = engine.MdxNet(mdx_weights.flatten_state(w, **hp), dim_t=cfg.dim_t, max_batch=BATCH, precision=1, **hp)
x = torch.randn(BATCH, 4, cfg.dim_t, cfg.dim_f, device="cuda")
for _ in range(2):
    net.forward(x, engine.LAYOUT_CTF)
torch.cuda.synchronize()
net.profile(True)
net.forward(x, engine.LAYOUT_CTF)
torch.cuda.synchronize()
pr = net.profile_read()
print("DBG=%s " % os.environ.get("B200SEP_DBG", "0") + " ".join(f"{k}={v['ms']:.2f}" for k, v in pr.items() if v["ms"] > 0) + f" total={sum(v['ms'] for v in pr.values()):.2f} ms")
