"""Real-time factor of the MDX23C / HTDemucs / VR paths on one B200 (measurement tool; bench.py stays the MDX north-star contract).

    python tests/dev/bench_arch.py [--arch mdxc|roformer|demucs|vr|all] [--steps 2] [--warmup 1] [--no-cpu]

One JSON line per architecture: value = audio-seconds / device-seconds with the mix resident in HBM (CUDA events),
e2e = through the engine's host-buffer entry (upload + download inside the timed region), cpu_baseline = the oracle port on the host
for a bounded sample (one forward).  Weights: seeded synthetic tensors of the released geometries (no model files offline)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "python-audio-separator_b200"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import mdx_oracle as M
from audio_separator.separator.b200._lib import launch_count

SR = 44100


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps / 1e3, (launch_count() - l0) // steps


def line(arch, workload, seconds, dev_s, e2e_s, launches, steps, warmup, h2d, d2h, cpu, extra):
    out = {"metric": "real-time factor (audio-sec/wall-sec) @44.1kHz stereo", "arch": arch, "value": round(seconds / dev_s, 1), "unit": "x realtime", "n_gpus": 1,
           "steps": steps, "warmup": warmup, "ms_per_step": round(dev_s * 1e3, 2), "higher_is_better": True, "dtype": "f32 (bf16x3 tensor-core contractions, fp32 accumulate)",
           "data": "synthetic", "config": {"workload": workload, **extra}, "e2e": {"value": round(seconds / e2e_s, 1), "unit": "x realtime", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h}, "gpu_launches": launches, "cpu_baseline": cpu}
    print(json.dumps(out), flush=True)


def bench_mdxc(a):
    import mdxc_oracle as X
    from audio_separator.separator.b200 import engine
    cfg = X.MDXCConfig()
    w = X.make_weights(cfg, seed=1, out_gain=0.3)
    net = engine.TfcNet(w, cfg.dim_f, cfg.dim_t, cfg.num_subbands, 2, cfg.num_scales, cfg.num_blocks_per_scale, cfg.num_channels_model, cfg.growth, cfg.bottleneck_factor, cfg.num_targets, max_batch=2)
    eng = engine.MdxcEngine(net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.dim_t, cfg.overlap)
    secs = a.minutes * 60
    mix = M.normalize(M.synth_music(int(secs * SR), seed=2), 0.9, 0.0)
    md = torch.from_numpy(mix).cuda()
    dev_s, launches = timed(lambda: eng.demix_device(md), a.steps, a.warmup)
    pin = torch.from_numpy(mix).pin_memory()
    e2e_s, _ = timed(lambda: eng.demix_device(pin.cuda(non_blocking=True)).cpu(), a.steps, 1)
    cpu = None
    if not a.no_cpu:
        t0 = time.time()
        X.net_forward(w, cfg, mix[None, :, : cfg.chunk_size])
        dt = time.time() - t0
        cpu = {"value": round(cfg.hop_size / SR / dt, 4), "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": "one chunk forward (hop_size new audio-seconds per chunk at overlap 8)"}
    line("MDXC", f"MDX23C-8KFFT-InstVoc_HQ topology, {a.minutes}-min track, overlap {cfg.overlap}, dim_t {cfg.dim_t}", secs, dev_s, e2e_s, launches, a.steps, a.warmup, mix.nbytes, 2 * mix.nbytes, cpu,
         {"chunks": X.chunk_grid(mix.shape[1], cfg)[3], "batch": 2})


def bench_roformer(a):
    import roformer_oracle as R
    from audio_separator.separator.b200 import roformer as rf
    kw = dict(stft_hop_length=441)
    ocfg = R.BSRoformerConfig(**kw)
    w = R.make_weights(ocfg, seed=8)
    eng = rf.RoformerEngine(rf.BSRoformerNet(rf.BSRoformerConfig(**kw), w), 801, 8, SR, n_instruments=2, batch_size=2)
    secs = a.minutes * 60
    mix = M.normalize(M.synth_music(int(secs * SR), seed=5), 0.9, 0.0)
    md = torch.from_numpy(mix).cuda()
    dev_s, launches = timed(lambda: eng.demix_device(md), a.steps, a.warmup)
    pin = torch.from_numpy(mix).pin_memory()
    e2e_s, _ = timed(lambda: eng.demix_device(pin.cuda(non_blocking=True)).cpu(), a.steps, 1)
    cpu = None
    if not a.no_cpu:
        t0 = time.time()
        R.forward(w, ocfg, mix[None, :, : ocfg.chunk_size])
        dt = time.time() - t0
        cpu = {"value": round((ocfg.chunk_size / SR) / dt, 4), "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": f"one 8-s chunk forward ({dt:.2f} s); step = chunk (no overlap at the default overlap setting)"}
    line("MDXC/BS-Roformer", f"model_bs_roformer_ep_317 geometry (dim 512, depth 12, 62 bands, n_fft 2048 / hop 441, dim_t 801), {a.minutes}-min track", secs, dev_s, e2e_s, launches, a.steps,
         a.warmup, mix.nbytes, mix.nbytes, cpu, {"chunks_per_forward": 2})


def bench_melband(a):
    import roformer_oracle as R
    from audio_separator.separator.b200 import roformer as rf
    ocfg = R.MelBandRoformerConfig()
    w = R.make_mel_weights(ocfg, seed=14)
    eng = rf.RoformerEngine(rf.BSRoformerNet(rf.MelBandRoformerConfig(stft_hop_length=441, mask_estimator_depth=2), w), 801, 8, SR, n_instruments=2, batch_size=2)
    secs = a.minutes * 60
    mix = M.normalize(M.synth_music(int(secs * SR), seed=6), 0.9, 0.0)
    md = torch.from_numpy(mix).cuda()
    dev_s, launches = timed(lambda: eng.demix_device(md), a.steps, a.warmup)
    pin = torch.from_numpy(mix).pin_memory()
    e2e_s, _ = timed(lambda: eng.demix_device(pin.cuda(non_blocking=True)).cpu(), a.steps, 1)
    cpu = None
    if not a.no_cpu:
        t0 = time.time()
        R.forward_mel(w, ocfg, mix[None, :, : ocfg.chunk_size])
        dt = time.time() - t0
        cpu = {"value": round((ocfg.chunk_size / SR) / dt, 4), "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": f"one 8-s chunk forward ({dt:.2f} s)"}
    line("MDXC/Mel-Band-Roformer", f"Mel-Band Roformer geometry (dim 384, depth 6, 60 mel bands, n_fft 2048 / hop 441, dim_t 801), {a.minutes}-min track", secs, dev_s, e2e_s, launches, a.steps,
         a.warmup, mix.nbytes, mix.nbytes, cpu, {"chunks_per_forward": 2})


def bench_demucs(a):
    import demucs_oracle as D
    from audio_separator.separator.b200 import demucs as dm
    ocfg = D.HTConfig()
    w = D.make_weights(ocfg, seed=11)
    eng = dm.DemucsEngine([dm.HTDemucsNet(dm.HTDemucsConfig(), w)], batch_size=4)
    secs = a.minutes * 60
    mix = M.normalize(M.synth_music(int(secs * SR), seed=3), 0.9, 0.0)
    offs = [[12623, 3268]]  # shifts = 2 (the plugin default), fixed draws
    md = torch.from_numpy(mix).cuda()
    dev_s, launches = timed(lambda: eng.apply_model(md, offs[0]), a.steps, a.warmup)
    e2e_s, _ = timed(lambda: eng.demix(mix, offs), a.steps, 1)
    cpu = None
    if not a.no_cpu:
        seg = mix[None, :, : ocfg.seg_len]
        t0 = time.time()
        D.forward(w, ocfg, seg)
        dt = time.time() - t0
        n_fwd = 2 * len(range(0, mix.shape[1] + 22050, int(0.75 * ocfg.seg_len)))
        cpu = {"value": round(secs / (dt * n_fwd), 4), "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": f"one segment forward ({dt:.2f} s) x {n_fwd} forwards per track"}
    line("Demucs", f"htdemucs geometry (1 model, 48 ch, 5-layer 512-d cross-transformer), {a.minutes}-min track, shifts 2, overlap 0.25, segment 7.8 s", secs, dev_s, e2e_s, launches, a.steps,
         a.warmup, mix.nbytes, 4 * mix.nbytes, cpu, {"segments_per_forward": 4})


def bench_vr(a):
    import vr_oracle as V
    from audio_separator.separator.b200 import vr
    arch = 537238
    w = V.make_weights(arch, seed=9)
    p4 = V.four_band_v2_param()
    eng = vr.VREngine(vr.VRNet(arch, 1344, w), p4, window_size=512, aggression=5, batch_size=4)
    secs = a.minutes * 60
    mix = M.normalize(M.synth_music(int(secs * SR), seed=4), 0.9, 0.0)
    md = torch.from_numpy(mix).cuda()

    def dev_step():
        spec = eng.loading_mix(md)
        y, v = eng.inference(spec)
        return eng.spec_to_wav(y), eng.spec_to_wav(v)

    dev_s, launches = timed(dev_step, a.steps, a.warmup)
    e2e_s, _ = timed(lambda: eng.separate(mix), a.steps, 1)
    cpu = None
    if not a.no_cpu:
        cfg = V.VRConfig(param=p4, nn_architecture=arch)
        x = np.abs(np.random.default_rng(0).standard_normal((1, 2, 673, 512))).astype(np.float32)
        t0 = time.time()
        V.predict_mask(w, cfg, x)
        dt = time.time() - t0
        cpu = {"value": round((256 * 480 / SR) / dt, 4), "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": f"one patch forward ({dt:.2f} s) = 256 frames x 480 samples of new audio"}
    line("VR", f"HP2 capacity CascadedASPPNet (537238), 4band_v2, {a.minutes}-min track, window 512, aggression 5", secs, dev_s, e2e_s, launches, a.steps, a.warmup, mix.nbytes, 2 * mix.nbytes, cpu,
         {"patches_per_forward": 4})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="all")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    for name, fn in (("mdxc", bench_mdxc), ("roformer", bench_roformer), ("melband", bench_melband), ("demucs", bench_demucs), ("vr", bench_vr)):
        if a.arch in ("all", name):
            fn(a)
