"""bench.py -- real-time factor of the MDX hot path (BASELINE.json metric) on N B200s, or the CPU reference arm.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--minutes 5]

Workload (BASELINE.json configs[1]): UVR-MDX-NET-Inst_HQ_3 topology (ConvTDFNet g=48, dim_f 3072, n_fft 6144),
5-minute 44.1 kHz stereo synthetic track, segment_size=256, overlap=0.25 -> 68 chunks per step.  There are no
model files offline, so the weights are seeded synthetic tensors of that architecture (data: "synthetic").

One step = separate the whole track: pad -> STFT -> net -> iSTFT -> Hann overlap-add -> normalise -> stems.
  value : audio-seconds / device-seconds with the track already resident in HBM (CUDA events, max over ranks)
  e2e   : same through the plugin-level API with the mix in pinned HOST memory and both float32 stems read back
N > 1  : chunks of the ONE track are time-sharded across ranks with an overlap-region halo exchange and a gather
         over NCCL (strong scaling), see audio_separator/separator/b200/sharded.py.
--impl reference: the reference's algorithm on the host CPU (torch-CPU restatement in oracle/, the reference
         package itself cannot be installed offline: no onnxruntime/librosa wheels) on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-audio-separator_b200"))

SR = 44100
# dram__bytes_read.sum + dram__bytes_write.sum of ONE conv3x3 launch at U-Net scale 0, batch 4 (profiles/r01b_conv3x3_s0_ncu_full.txt)
NCU_CONV_S0_DRAM_BYTES_PER_LAUNCH = 731.8e6 + 605.5e6
METRIC = "real-time factor (audio-sec/wall-sec) @44.1kHz stereo"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "source": "measured (MEASURED_PEAKS.json, sustained bf16 / copy)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons), "samples": len(self.rows)}


_CPU_THREADS = None


def _pick_cpu_threads(O, cfg, w):
    """All host threads the CPU path can use PRODUCTIVELY: torch-CPU convolutions stop scaling (and regress) well before 128
    threads, so time one network forward at a few thread counts and keep the fastest (reported as `cores`)."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import numpy as np
    import torch

    n = os.cpu_count() or 1
    cands = sorted({n, max(1, n // 2), min(n, 32), min(n, 16)}, reverse=True)
    x = np.random.default_rng(0).standard_normal((1, 4, cfg.dim_f, cfg.dim_t)).astype(np.float32)
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.perf_counter()
        O.convtdfnet_forward(w, cfg, x)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    _CPU_THREADS = best
    torch.set_num_threads(best)
    return best


def cpu_reference_rtf(sample_seconds, cfg_kwargs=None, repeats=1):
    """The reference's algorithm on the host cores: oracle.demix with the torch-CPU ConvTDFNet (fp32).
    Returns (rtf, seconds_of_audio, wall, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mdx_oracle as O

    cfg = O.MDXConfig(**(cfg_kwargs or {}))
    n = int(sample_seconds * SR)
    mix = O.normalize(O.synth_music(n, seed=1234), 0.9, 0.0)
    w = O.make_convtdfnet_weights(cfg, seed=11)
    cores = _pick_cpu_threads(O, cfg, w)
    t0 = time.perf_counter()
    for _ in range(repeats):
        O.demix(mix, cfg, lambda s: O.convtdfnet_forward(w, cfg, s))
    wall = (time.perf_counter() - t0) / repeats
    return n / SR / wall, n / SR, wall, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 10.0  # BASELINE configs[0]: 10 s = 3 chunks of the same grid
    # warm-up = the thread-count calibration inside the first call (pages in MKL/oneDNN); each further pass is 3 forwards
    walls = []
    for _ in range(args.steps):
        rtf, secs, wall, cores = cpu_reference_rtf(sample)
        walls.append(wall)
    wall = sum(walls) / len(walls)
    value = sample / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UVR-MDX-NET-Inst_HQ_3 topology, 44.1 kHz stereo synthetic, segment_size=256, overlap=0.25", "sample": "10 s excerpt (3 chunks) per step"},
        "cpu_baseline": {"value": value, "unit": "x realtime", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port", "sample": "10 s excerpt = 3 chunks of the 68-chunk grid, torch-CPU fp32 ConvTDFNet + numpy STFT/OLA (oracle/mdx_oracle.py); reference package not installable offline (onnxruntime, librosa wheels absent)"},
        "e2e": {"value": value, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--minutes", type=float, default=5.0)
    ap.add_argument("--batch", type=int, default=4, help="chunks per network forward")
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mdx_oracle as O  # synthetic weights + programme material generators (and the cpu_baseline leg below)
    from audio_separator.separator.b200 import _lib, engine, mdx_weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}"

    cfg = O.MDXConfig()
    N = int(args.minutes * 60 * SR)
    audio_seconds = N / SR
    mix = O.synth_music(min(N, 30 * SR), seed=1234)
    reps = -(-N // mix.shape[1])
    mix = np.tile(mix, (1, reps))[:, :N].copy()  # 30 s pattern tiled: > L2 by far (106 MB), deterministic
    w = O.make_convtdfnet_weights(cfg, seed=11, out_gain=0.02)
    hp = mdx_weights.infer_hparams_from_state(w)
    net = engine.MdxNet(mdx_weights.flatten_state(w, **hp), dim_t=cfg.dim_t, max_batch=args.batch, precision=args.precision, **hp)
    if world > 1:
        from audio_separator.separator.b200.sharded import ShardedMdxEngine

        eng = ShardedMdxEngine(net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.segment_size, cfg.overlap, cfg.compensate, batch_size=args.batch)
    else:
        eng = engine.MdxEngine(net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.segment_size, cfg.overlap, cfg.compensate, batch_size=args.batch)

    mix_host = torch.from_numpy(mix).pin_memory()
    mix_dev = mix_host.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return eng.separate_device(mix_dev, 0.9, 0.0)

    out_host = [torch.empty((N, 2), dtype=torch.float32).pin_memory() for _ in range(2)]

    def step_e2e():
        d = mix_host.cuda(non_blocking=True)
        p, s = eng.separate_device(d, 0.9, 0.0)
        if p is not None:
            out_host[0].copy_(p, non_blocking=True)
            out_host[1].copy_(s, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.summary() if rank == 0 else None

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 * 0.0)

    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])

    # ---- roofline of the dominant kernel (launch shape), timed live with CUDA events around each launch on the engine's stream
    net.profile(True)
    step_device()
    torch.cuda.synchronize()
    prof = net.profile_read()
    net.profile(False)
    peaks = load_peaks()
    net_ms = sum(v["ms"] for v in prof.values())
    # dominant kernel = the single launch SHAPE with the most device time: the 3x3 convolutions at U-Net scale 0 (one shape, 2 launches per block);
    # the "conv3x3" category aggregates the five deeper, smaller shapes
    tname, tv = ("conv3x3_scale0", prof["conv3x3_scale0"]) if prof.get("conv3x3_scale0", {}).get("ms", 0) > 0 else max(prof.items(), key=lambda kv: kv[1]["ms"])
    tf = tv["flops"] / (tv["ms"] * 1e-3) / 1e12 if tv["ms"] > 0 else 0.0
    # DRAM bytes of that launch shape from the committed `ncu --set full` capture (profiles/README.md); only valid for the default config
    traffic = None
    if tname == "conv3x3_scale0" and args.batch == 4:
        traffic = NCU_CONV_S0_DRAM_BYTES_PER_LAUNCH
    roofline = {
        "kernel": f"umma_pair_kernel[{tname}]" if args.precision else f"conv2d_simt_kernel[{tname}]", "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"],
        "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"], "traffic": traffic, "peak_source": peaks["source"],
        "arithmetic": "bf16x3 split (3 tcgen05 MMAs per algorithmic MAC): attainable = peak/3" if args.precision else "fp32 FMA (no tensor cores)",
        "frac_of_attainable": (3.0 if args.precision else 1.0) * tf / peaks["bf16_tflops"],
        "launches": tv["launches"], "avg_launch_ms": tv["ms"] / max(1, tv["launches"]),
        "algorithmic_flops_per_launch": tv["flops"] / max(1, tv["launches"]), "algorithmic_bytes_per_launch": tv["bytes"] / max(1, tv["launches"]),
        "hbm_view": {"achieved_gbs": tv["bytes"] / (tv["ms"] * 1e-3) / 1e9 if tv["ms"] > 0 else 0.0, "peak_gbs": peaks["hbm_gbs"]},
        "share_of_net_time": tv["ms"] / max(1e-9, net_ms), "net_ms_per_step": net_ms,
        "by_category_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]},
        "by_category_tflops": {k: round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) for k, v in prof.items() if v["ms"] > 0 and v["flops"] > 0},
    }

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    value = audio_seconds / (ms_step * 1e-3)
    e2e_val = audio_seconds / (ms_e2e / args.steps * 1e-3)
    cpu_b = None
    if not args.no_cpu_baseline:
        rtf, secs, wall, cores = cpu_reference_rtf(10.0)
        cpu_b = {"value": rtf, "unit": "x realtime", "cores": cores, "kind": "port", "sample": f"10 s excerpt (3 of 68 chunks) of the same workload, {wall:.1f} s CPU wall, torch-CPU fp32 net + numpy STFT/OLA (oracle/mdx_oracle.py)"}
    L, step, n_chunks, _ = engine.MdxEngine.grid(eng, N) if hasattr(eng, "grid") else (0, 0, 0, 0)
    line = {
        "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"UVR-MDX-NET-Inst_HQ_3 topology (ConvTDFNet g=48, dim_f=3072, n_fft=6144), {args.minutes:g}-min 44.1 kHz stereo synthetic, segment_size=256, overlap=0.25, {n_chunks} chunks/step",
                   "batch": args.batch, "precision": args.precision, "l2": "inputs larger than L2 (106 MB track, 0.6-4.8 GB activations per forward)", "parallelism": f"time-sharded chunks x{args.gpus}" if args.gpus > 1 else "single GPU"},
        "e2e": {"value": e2e_val, "unit": "x realtime", "h2d_bytes_per_step": int(mix_host.numel() * 4), "d2h_bytes_per_step": int(2 * N * 2 * 4)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_b,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
