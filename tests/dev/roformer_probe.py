"""Full-size BS-Roformer chunk timing on the GPU box (dev tool): python tests/dev/roformer_probe.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import roformer_oracle as R
from audio_separator.separator.b200 import roformer as rf
from audio_separator.separator.b200._lib import launch_count

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kw = dict(stft_hop_length=441)
ocfg = R.BSRoformerConfig(**kw)
net = rf.BSRoformerNet(rf.BSRoformerConfig(**kw), R.make_weights(ocfg, seed=8))
x = torch.randn((B, 2, ocfg.chunk_size), device="cuda") * 0.2
for _ in range(2):
    net.forward(x)
torch.cuda.synchronize()
l0 = launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    net.forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"batch {B}: {ms:.1f} ms / forward ({ms / B:.1f} per 8-s chunk), launches {(launch_count() - l0) // n}, RTF ~ {8.0 * B / (ms / 1e3):.0f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.forward(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14))
