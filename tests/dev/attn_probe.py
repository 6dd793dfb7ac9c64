"""The batched attention GEMMs of the BS-Roformer time transformer in isolation (dev tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200")]
import torch
from audio_separator.separator.b200._lib import check, lib

B, n, dh, ldv = 496, 801, 64, 804
q, k = torch.randn((B, n, dh), device="cuda"), torch.randn((B, n, dh), device="cuda")
vt = torch.randn((B, dh, ldv), device="cuda")
sc = torch.empty((B, n, ldv), device="cuda")
o = torch.empty((B, n, dh), device="cuda")


def scores():
    check(lib.b200sep_gemm_f32(q.data_ptr(), k.data_ptr(), sc.data_ptr(), n, n, dh, dh, dh, ldv, B, n * dh, n * dh, n * ldv, 0.125, None, None, 0, None, None, None, 0))


def pv():
    check(lib.b200sep_gemm_f32(sc.data_ptr(), vt.data_ptr(), o.data_ptr(), n, dh, n, ldv, ldv, dh, B, n * ldv, dh * ldv, n * dh, 1.0, None, None, 0, None, None, None, 0))


for name, fn, fl in (("scores", scores, 2 * B * n * n * dh), ("pv", pv, 2 * B * n * n * dh)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms * 1e3:.0f} us  {fl / ms / 1e9:.1f} TFLOP/s")
