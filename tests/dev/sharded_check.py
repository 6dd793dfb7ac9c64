"""Development check (torchrun, >= 2 GPUs): ShardedMdxEngine == MdxEngine on the same input, bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("python-audio-separator_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch, torch.distributed as dist
import mdx_oracle as O
from audio_separator.separator.b200 import engine, mdx_weights
from audio_separator.separator.b200.sharded import ShardedMdxEngine

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
cfg = O.MDXConfig(n_fft=1024, hop_length=256, dim_f=512, dim_t=32, segment_size=32, g=16, num_blocks=7)
w = O.make_convtdfnet_weights(cfg, seed=21, out_gain=0.05)
hp = mdx_weights.infer_hparams_from_state(w)
net = engine.MdxNet(mdx_weights.flatten_state(w, **hp), dim_t=cfg.dim_t, max_batch=3, precision=1, **hp)
mix = torch.as_tensor(O.synth_music(200_000, seed=3)).cuda()
args = (net, cfg.n_fft, cfg.hop_length, cfg.dim_f, cfg.segment_size, cfg.overlap, cfg.compensate)
p1, s1 = engine.MdxEngine(*args).separate_device(mix)
p2, s2 = ShardedMdxEngine(*args).separate_device(mix)
torch.cuda.synchronize()
if rank == 0:
    ok = torch.equal(p1, p2) and torch.equal(s1, s2)
    print(f"sharded x{world}: identical to single-GPU = {ok}; max|diff| = {(p1 - p2).abs().max().item():.3e}")
    ref_p, ref_s = O.separate_arrays(mix.cpu().numpy(), cfg, lambda s: O.convtdfnet_forward(w, cfg, s))
    print(f"vs oracle: primary {np.abs(p2.cpu().numpy() - ref_p).max():.2e} secondary {np.abs(s2.cpu().numpy() - ref_s).max():.2e}")
    assert ok
dist.barrier(); dist.destroy_process_group()
