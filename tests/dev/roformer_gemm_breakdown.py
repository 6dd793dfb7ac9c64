"""Per-shape time of the GEMMs inside one full-size BS-Roformer forward (dev tool; every call is timed with events and synchronised)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200"), os.path.join(ROOT, "oracle")]
import torch
import roformer_oracle as R
from audio_separator.separator.b200 import roformer as rf
from audio_separator.separator.b200 import _lib

kw = dict(stft_hop_length=441)
ocfg = R.BSRoformerConfig(**kw)
net = rf.BSRoformerNet(rf.BSRoformerConfig(**kw), R.make_weights(ocfg, seed=8))
x = torch.randn((1, 2, ocfg.chunk_size), device="cuda") * 0.2
net.forward(x)
torch.cuda.synchronize()
stats = collections.defaultdict(lambda: [0, 0.0])
orig = _lib.lib.b200sep_gemm_f32


def timed(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = orig(*a)
    e1.record()
    torch.cuda.synchronize()
    M, N, K, batch = a[3], a[4], a[5], a[9]
    key = (M, N, K, batch)
    stats[key][0] += 1
    stats[key][1] += e0.elapsed_time(e1)
    return rc


_lib.lib.b200sep_gemm_f32 = timed
rf.lib.b200sep_gemm_f32 = timed
net.forward(x)
tot = sum(v[1] for v in stats.values())
print(f"total GEMM time {tot:.1f} ms")
for k, (n, ms) in sorted(stats.items(), key=lambda t: -t[1][1])[:16]:
    M, N, K, b = k
    print(f"M={M:6d} N={N:5d} K={K:5d} batch={b:5d}: {n:3d} calls {ms:7.2f} ms  {2 * M * N * K * b * n / ms / 1e9:7.1f} TFLOP/s")
