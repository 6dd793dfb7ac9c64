"""Full-size HTDemucs forward timing on the GPU box (dev tool): python tests/dev/demucs_probe.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import demucs_oracle as D
from audio_separator.separator.b200 import demucs as dm
from audio_separator.separator.b200._lib import launch_count

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ocfg = D.HTConfig()
w = D.make_weights(ocfg, seed=11)
net = dm.HTDemucsNet(dm.HTDemucsConfig(), w)
x = torch.randn((B, 2, ocfg.seg_len), device="cuda") * 0.3
if os.environ.get("ONCE"):  # under ncu: exactly one forward
    net.forward(x)
    torch.cuda.synchronize()
    sys.exit(0)
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
print(f"batch {B}: {ms:.1f} ms / forward ({ms / B:.1f} per segment), wall {(time.time() - t0) / n * 1e3:.1f} ms, launches {(launch_count() - l0) // n}, "
      f"RTF(one model, shifts 0) ~ {0.75 * 7.8 * B / (ms / 1e3):.1f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.forward(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))
