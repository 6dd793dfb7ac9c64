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
    # segments_enabled=False: apply_model(split=False) -- one forward over the whole (shifted) track (apply.py:251-260)
    random.seed(4)
    state = random.getstate()
    with torch.no_grad():
        ns_ref = apply.apply_model(model, mn[None], shifts=1, split=False, device="cpu").numpy()
    random.setstate(state)
    ns_offs = [random.randint(0, int(0.5 * cfg.samplerate))]

    def whole(fn_, c_, mixn, offs_, valid_of):  # restated: pad by max_shift, TensorChunk(offset, length + max_shift - offset), centred zero-padding to valid_length, center_trim
        import demucs_oracle as D

        ms = int(0.5 * c_.samplerate)
        Nn = mixn.shape[-1]
        pm = D.padded(mixn, 0, Nn, Nn + 2 * ms)
        out = 0
        for o in offs_:
            ln = Nn + ms - o
            y = D.center_trim(fn_(D.padded(pm, o, ln, valid_of(ln))), ln)
            out = out + y[..., ms - o :]
        return out / len(offs_)

    check("apply_model shifts=1 split=False (HDemucs: no valid_length)", ns_ref, whole(fn, cfg, mn.numpy()[None], ns_offs, lambda ln: ln), 5e-5 * max(1.0, np.abs(ns_ref).max()))
    # the same for HTDemucs on a clip shorter than its training segment (valid_length = the training segment; longer inputs raise, htdemucs.py:469-481)
    import demucs_oracle as D
    from make_golden_demucs import SMALL as HT_SMALL, ref_model as ht_ref_model

    hcfg = D.HTConfig(**HT_SMALL)
    hw = D.make_weights(hcfg, seed=5)
    hmodel = ht_ref_model(hcfg, hw)
    hN = hcfg.seg_len - 1000  # shifts = 0: the small model's training segment (0.5 s) is no longer than max_shift, so a shifted chunk never fits
    hm = torch.from_numpy(mix[:, :hN])
    hmn = (hm - hm.mean(0).mean()) / hm.mean(0).std()
    hoffs = []
    with torch.no_grad():
        hns_ref = apply.apply_model(hmodel, hmn[None], shifts=0, split=False, device="cpu").numpy()

    def whole0(fn_, mixn, valid):
        return D.center_trim(fn_(D.padded(mixn, 0, mixn.shape[-1], valid)), mixn.shape[-1])

    check("apply_model shifts=0 split=False (HTDemucs, padded to the training segment)", hns_ref,
          whole0(lambda c: D.forward(hw, hcfg, c), hmn.numpy()[None], hcfg.seg_len), 5e-5 * max(1.0, np.abs(hns_ref).max()))
    try:
        with torch.no_grad():
            apply.apply_model(hmodel, torch.zeros(1, 2, hcfg.seg_len + 10), shifts=0, split=False, device="cpu")
        raise AssertionError("the reference accepted an input longer than the training segment")
    except ValueError as e:
        print("  reference refuses long inputs without segments:", e)
    np.savez_compressed(
        os.path.join(GOLD, "hdemucs_small.npz"), nosplit_ref=ns_ref.astype(np.float32), nosplit_offsets=np.array(ns_offs), ht_nosplit_ref=hns_ref.astype(np.float32),
        ht_nosplit_offsets=np.array(hoffs), ht_nosplit_n=hN, weights_seed=9, mix_seed=41, seg_len=L, n_apply=N, shift_offsets=np.array(offs), long_len=long.shape[-1],
        forward_ref=y_ref.astype(np.float32), forward_long_ref=yl_ref.astype(np.float32), forward_old_ref=yo_ref.astype(np.float32), forward_b2_ref=y2_ref.astype(np.float32),
        apply_ref=a_ref.astype(np.float32), apply0_ref=b_ref.astype(np.float32),
    )
    print("wrote tests/golden/hdemucs_small.npz; oracle pinned: OK")


if __name__ == "__main__":
    main()
