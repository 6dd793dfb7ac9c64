"""Pin oracle/demucs_oracle.py against the UNMODIFIED reference HTDemucs + apply_model and write tests/golden/demucs_small.npz."""
import os
import random
import sys
from fractions import Fraction

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import demucs_oracle as D  # noqa: E402
import mdx_oracle as M  # noqa: E402
import ref_shim  # noqa: E402
from make_golden import check  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
SMALL = dict(channels=8, bottom_channels=32, t_layers=3, t_heads=4, segment=Fraction(1, 2))


def ref_model(cfg, w):
    ht = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.demucs.htdemucs")
    m = ht.HTDemucs(**cfg.kwargs()).eval()
    sd = m.state_dict()
    assert [k for k in sd] == [n for n, _ in D.param_shapes(cfg)], [(a, b) for a, (b, _) in zip(sd, D.param_shapes(cfg)) if a != b][:5]
    for (n, s), v in zip(D.param_shapes(cfg), sd.values()):
        assert tuple(v.shape) == s, (n, v.shape, s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return m


def main():
    cfg = D.HTConfig(**SMALL)
    w = D.make_weights(cfg, seed=5)
    model = ref_model(cfg, w)
    L = cfg.seg_len
    mix = M.synth_music(3 * L, seed=31)
    seg = mix[None, :, :L]
    with torch.no_grad():
        y_ref = model(torch.from_numpy(seg)).numpy()
    y_orc = D.forward(w, cfg, seg)
    check("htdemucs forward (full segment)", y_ref, y_orc, 2e-5 * max(1.0, np.abs(y_ref).max()))
    short = mix[None, :, : L - 1234]
    with torch.no_grad():
        ys_ref = model(torch.from_numpy(short)).numpy()
    check("htdemucs forward (short, padded)", ys_ref, D.forward(w, cfg, short), 2e-5 * max(1.0, np.abs(ys_ref).max()))

    # constructor defaults that differ from the pinned structure: DConv in the encoders only (dconv_mode=1), no channel up-samplers
    cfg1 = D.HTConfig(**dict(SMALL, dconv_mode=1, bottom_channels=0, t_layers=2))
    w1 = D.make_weights(cfg1, seed=6)
    model1 = ref_model(cfg1, w1)
    with torch.no_grad():
        y1_ref = model1(torch.from_numpy(seg)).numpy()
    check("htdemucs forward (dconv_mode=1, bottom_channels=0)", y1_ref, D.forward(w1, cfg1, seg), 2e-5 * max(1.0, np.abs(y1_ref).max()))

    # apply_model with shifts=2, split=True: record the reference's random offsets
    apply = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.demucs.apply")
    N = int(2.3 * L)
    m2 = mix[:, :N]
    ref = torch.from_numpy(m2).mean(0)
    mn = (torch.from_numpy(m2) - ref.mean()) / ref.std()
    random.seed(0)
    # the shift offsets come from Python's global RNG, which the transformer's positional embedding also draws from inside every
    # forward (transformer.py:562) -> record the values apply_model actually receives instead of re-deriving them
    offs = []
    real_randint = apply.random.randint

    class _Rec:
        def __getattr__(self, name):
            return getattr(random, name)

        def randint(self, a, b):
            v = real_randint(a, b)
            offs.append(v)
            return v

    apply.random = _Rec()
    with torch.no_grad():
        a_ref = apply.apply_model(model, mn[None], shifts=2, split=True, overlap=0.25, device="cpu").numpy()
    apply.random = random
    print("  recorded shift offsets:", offs)
    fn = lambda c: D.forward(w, cfg, c)  # noqa: E731
    a_orc = D.apply_model(fn, cfg, mn.numpy()[None], offs, 0.25)
    check("apply_model shifts=2 split", a_ref, a_orc, 5e-5 * max(1.0, np.abs(a_ref).max()))
    with torch.no_grad():
        b_ref = apply.apply_model(model, mn[None], shifts=0, split=True, overlap=0.25, device="cpu").numpy()
    check("apply_model shifts=0 split", b_ref, D.apply_model(fn, cfg, mn.numpy()[None], [], 0.25), 5e-5 * max(1.0, np.abs(b_ref).max()))
    # demix_demucs glue (demucs_separator.py:162-195) restated on top of the reference's apply_model
    src_ref = (torch.from_numpy(a_ref[0]) * ref.std() + ref.mean()).numpy()
    src_ref[[0, 1]] = src_ref[[1, 0]]
    src_orc = D.demix_demucs([fn], [[1.0] * 4], cfg, m2, [offs], 0.25)
    check("demix_demucs", src_ref, src_orc, 5e-5 * max(1.0, np.abs(src_ref).max()))
    np.savez_compressed(
        os.path.join(GOLD, "demucs_small.npz"), weights_seed=5, mix_seed=31, seg_len=L, n_apply=N, shift_offsets=np.array(offs),
        forward_ref=y_ref.astype(np.float32), forward_short_ref=ys_ref.astype(np.float32), forward_dm1_ref=y1_ref.astype(np.float32), apply_ref=a_ref.astype(np.float32), demix_ref=src_ref.astype(np.float32),
    )
    print("wrote tests/golden/demucs_small.npz; oracle pinned: OK")


if __name__ == "__main__":
    main()
