"""Throughput of the fp32 tensor-core GEMM / conv on representative shapes (dev tool).  B200SEP_TC=0 times the SIMT kernels instead."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200")]
import numpy as np, torch
from audio_separator.separator.b200 import demucs as dm


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("TC", os.environ.get("B200SEP_TC", "1"))
for M, N, K in ((2688, 512, 512), (2688, 2048, 512), (2688, 512, 2048), (1344, 1536, 512), (10752, 2048, 512)):
    x, w, b = torch.randn((M, K), device="cuda"), torch.randn((N, K), device="cuda"), torch.randn(N, device="cuda")
    ms = timeit(lambda: dm.linear(x, w, b))
    print(f"linear M={M} N={N} K={K}: {ms * 1e3:.0f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
for (B, cin, cout, H, W, k, s) in ((1, 64, 64, 336, 512, 3, 1), (1, 128, 128, 672, 512, 3, 1), (1, 192, 128, 672, 512, 3, 1), (1, 512, 512, 84, 64, 3, 1), (1, 128, 128, 336, 256, 3, 2),
                                  (1, 3584, 1024, 42, 32, 1, 1), (4, 128, 128, 672, 512, 3, 1)):
    x = torch.randn((B, cin, H, W), device="cuda")
    wb = torch.from_numpy(dm.block_conv_weight(np.random.randn(cout, cin, k, k).astype(np.float32) / (cin * k * k) ** 0.5)).cuda()
    bias = torch.randn(cout, device="cuda")
    ms = timeit(lambda: dm.conv2d(x, wb, bias, cout, (k, k), s=(s, s), p=(k // 2, k // 2), act=1))
    Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
    print(f"conv B={B} {cin}->{cout} {H}x{W} k{k} s{s}: {ms * 1e3:.0f} us  {2 * B * Ho * Wo * cout * cin * k * k / ms / 1e9:.1f} TFLOP/s")
