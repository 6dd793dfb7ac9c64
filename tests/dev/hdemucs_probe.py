"""Full-size Hybrid Demucs v3 forward timing on the GPU box (dev tool): python tests/dev/hdemucs_probe.py [batch] [seconds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import hdemucs_oracle as H
from audio_separator.separator.b200 import hdemucs as hd
from audio_separator.separator.b200._lib import launch_count

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
ocfg = H.HDConfig()
w = H.make_weights(ocfg, seed=11)
net = hd.HDemucsNet(hd.HDemucsConfig(), w)
L = int(secs * 44100)
x = torch.randn((B, 2, L), device="cuda") * 0.3
for _ in range(2):
    y = net.forward(x)
torch.cuda.synchronize()
l0 = launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
n = 3
for _ in range(n):
    y = net.forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"batch {B} x {secs:.0f} s: {ms:.1f} ms / forward ({ms / B:.1f} per segment), wall {(time.time() - t0) / n * 1e3:.1f} ms, launches {(launch_count() - l0) // n}, "
      f"RTF(one model, shifts 0, overlap 0.25) ~ {0.75 * secs * B / (ms / 1e3):.1f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.forward(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
