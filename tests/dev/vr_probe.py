"""Full-size VR patch timing on the GPU box (dev tool): python tests/dev/vr_probe.py [arch] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import vr_oracle as V
from audio_separator.separator.b200 import vr
from audio_separator.separator.b200._lib import launch_count

arch = int(sys.argv[1]) if len(sys.argv) > 1 else 537238
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = vr.VRNet(arch, 1344, V.make_weights(arch, seed=9))
x = torch.rand((B, 2, 673, 512), device="cuda")
for _ in range(2):
    net.predict_mask(x)
torch.cuda.synchronize()
l0 = launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    net.predict_mask(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"arch {arch} batch {B}: {ms:.1f} ms / forward ({ms / B:.1f} per patch), launches {(launch_count() - l0) // n}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB; "
      f"3-min track (65 patches) ~ {65 * ms / B / 1e3:.2f} s -> RTF ~ {180 / (65 * ms / B / 1e3):.0f}")
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.predict_mask(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12))
