"""Pin oracle/hdemucs_oracle.py against the UNMODIFIED reference HDemucs (+ apply_model) and write tests/golden/hdemucs_small.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import hdemucs_oracle as H  # noqa: E402
import mdx_oracle as M  # noqa: E402
import ref_shim  # noqa: E402
from make_golden import check  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
# depth 4 at nfft 256: frequency layers 128 -> 32 -> 8 -> (last_freq) 1, then one time layer; GroupNorm / BLSTM / LocalState from index 2 on:
# the same layer kinds, in the same order, as the released depth-6 / nfft-4096 models (norm_starts = dconv_lstm = dconv_attn = 4)
SMALL = dict(channels=8, nfft=256, depth=4, norm_starts=2, dconv_lstm=2, dconv_attn=2, segment=0.5)


def ref_model(cfg, w):
    hd = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.demucs.hdemucs")
    m = hd.HDemucs(**cfg.kwargs()).eval()
    sd = m.state_dict()
    names = [n for n, _ in H.param_shapes(cfg)]
    assert list(sd) == names, [(a, b) for a, b in zip(sd, names) if a != b][:5] + [len(sd), len(names)]
    for (n, s), v in zip(H.param_shapes(cfg), sd.values()):
        assert tuple(v.shape) == s, (n, v.shape, s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return m


def main():
    cfg = H.HDConfig(**SMALL)
    w = H.make_weights(cfg, seed=9)
    model = ref_model(cfg, w)
    L = cfg.seg_len  # 22050 samples: 87 frames; the time branch reaches 87 steps at the merge
    mix = M.synth_music(3 * L, seed=41)
    seg = mix[None, :, :L]
    with torch.no_grad():
        y_ref = model(torch.from_numpy(seg)).numpy()
    y_orc = H.forward(w, cfg, seg)
    check("hdemucs forward (one segment)", y_ref, y_orc, 2e-5 * max(1.0, np.abs(y_ref).max()))
    # a length that is not a multiple of anything and long enough for the BLSTM to frame its input (T > 200 steps at the merge layer: > 51200 samples)
    long = mix[None, :, : 2 * L + 12345]
    with torch.no_grad():
        yl_ref = model(torch.from_numpy(long)).numpy()
    check("hdemucs forward (ragged length, framed BLSTM)", yl_ref, H.forward(w, cfg, long), 2e-5 * max(1.0, np.abs(yl_ref).max()))
    # hybrid_old (zero instead of reflect padding of the spectrogram input, hdemucs.py:537-541 / :558-567): the mdx_extra checkpoints
    cfg_o = H.HDConfig(**dict(SMALL, hybrid_old=True))
    model_o = ref_model(cfg_o, w)
    with torch.no_grad():
        yo_ref = model_o(torch.from_numpy(seg)).numpy()
    check("hdemucs forward (hybrid_old)", yo_ref, H.forward(w, cfg_o, seg), 2e-5 * max(1.0, np.abs(yo_ref).max()))
    # two batch entries
    two = np.stack([mix[:, :L], mix[:, L : 2 * L]])
    with torch.no_grad():
        y2_ref = model(torch.from_numpy(two)).numpy()
    check("hdemucs forward (batch 2)", y2_ref, H.forward(w, cfg, two), 2e-5 * max(1.0, np.abs(y2_ref).max()))

    # apply_model, split=True, shifts=0 and shifts=1 (HDemucs draws nothing from the global RNG inside forward)
    import random

    apply = ref_shim.ref_module("audio_separator.separator.uvr_lib_v5.demucs.apply")
    N = int(2.3 * L)
    m2 = mix[:, :N]
    ref = torch.from_numpy(m2).mean(0)
    mn = (torch.from_numpy(m2) - ref.mean()) / ref.std()
    fn = lambda c: H.forward(w, cfg, c)  # noqa: E731
    with torch.no_grad():
        b_ref = apply.apply_model(model, mn[None], shifts=0, split=True, overlap=0.25, device="cpu").numpy()
    check("apply_model shifts=0 split", b_ref, H.apply_model(fn, cfg, mn.numpy()[None], [], 0.25), 5e-5 * max(1.0, np.abs(b_ref).max()))
    random.seed(3)
    state = random.getstate()
    with torch.no_grad():
        a_ref = apply.apply_model(model, mn[None], shifts=1, split=True, overlap=0.25, device="cpu").numpy()
    random.setstate(state)
    offs = [random.randint(0, int(0.5 * cfg.samplerate))]
    print("  shift offsets:", offs)
    check("apply_model shifts=1 split", a_ref, H.apply_model(fn, cfg, mn.numpy()[None], offs, 0.25), 5e-5 * max(1.0, np.abs(a_ref).max()))
    np.savez_compressed(
        os.path.join(GOLD, "hdemucs_small.npz"), weights_seed=9, mix_seed=41, seg_len=L, n_apply=N, shift_offsets=np.array(offs), long_len=long.shape[-1],
        forward_ref=y_ref.astype(np.float32), forward_long_ref=yl_ref.astype(np.float32), forward_old_ref=yo_ref.astype(np.float32), forward_b2_ref=y2_ref.astype(np.float32),
        apply_ref=a_ref.astype(np.float32), apply0_ref=b_ref.astype(np.float32),
    )
    print("wrote tests/golden/hdemucs_small.npz; oracle pinned: OK")


if __name__ == "__main__":
    main()
