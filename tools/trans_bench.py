"""Throughput of the trans-dimensional path (BASELINE config 3: transepic, B=8192, N=128, bf16):
one TransdimensionalEPiC evaluation (mmb_trans_forward) against the bf16 tensor roofline
(144.5 MFLOP per jet-evaluation in the two transformer stacks + 0.83 in the trunk),
the fused sampler update against the HBM roofline (133 B per live particle-step with in-kernel Philox),
and a short JumpSampler run (dt = 0.02 -> 50 evaluations) as generated jets/s.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from multimodal_particles_b200 import _native  # noqa: E402
from multimodal_particles_b200.config_classes.transdimensional_unconditional_config import TransdimensionalEpicConfig  # noqa: E402
from multimodal_particles_b200.transdimensional import JumpSampler, TransdimensionalJumpDiffusion  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N, S = 128, 8
cfg = TransdimensionalEpicConfig()
torch.manual_seed(0)
model = TransdimensionalJumpDiffusion(cfg).to(dev)
m = model.net.model
trunk, heads = m.native_trunk(dev), m.native_heads(dev)
g = torch.Generator().manual_seed(1234)
dims = torch.randint(1, N + 1, (B,), generator=g)
mask = (torch.arange(N)[None] < dims[:, None]).float().unsqueeze(-1)
x = torch.randn(B, N, 3, generator=g) * mask
x = x - (x.sum(1, keepdim=True) / dims.view(B, 1, 1)) * mask
oh = torch.randn(B, N, S, generator=g) * mask
ts = torch.rand(B, generator=g) * 0.999 + 1e-3
near = (torch.rand(B, generator=g) * dims).long()
x, oh, dims32, ts, near = x.to(dev), oh.to(dev), dims.to(dev, torch.int32), ts.to(dev), near.to(dev, torch.int32)
fr = model.forward_rate.as_c()


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        out.append(s.elapsed_time(e))
    return sum(out) / len(out)


fwd_ms = timed(lambda: _native.trans_forward(trunk, heads, x, oh, dims32, ts, near, None, fr, precision="bf16", want_auto=False), 5)
pk = bench.peaks()
flops = (144.5e6 + 0.83e6) * B
tf = flops / (fwd_ms * 1e-3) / 1e12

# fused sampler update alone (Philox noise), 133 B per live particle-step, on a batch larger than L2 (32768 jets ~ 280 MB);
# ten launches per event pair so that the Python launch gap between the events does not count as kernel time
BU = 32768
gu = torch.Generator().manual_seed(7)
dims_u = torch.randint(1, N + 1, (BU,), generator=gu)
mask_u = (torch.arange(N)[None] < dims_u[:, None]).float().unsqueeze(-1).to(dev)
xs, ohs, ds = torch.randn(BU, N, 3, device=dev) * mask_u, torch.randn(BU, N, S, device=dev) * mask_u, dims_u.to(dev, torch.int32)
v, lg = torch.randn(BU, N, 3, device=dev), torch.randn(BU, N, S, device=dev)
rate = torch.zeros(BU, device=dev)          # no births: the live set (hence the byte count) stays fixed over the repetitions
nm, ns = torch.randn(BU, 3 + S, device=dev), torch.randn(BU, 3 + S, device=dev)
INNER = 10


def upd():
    for _ in range(INNER):
        _native.trans_sampler_update(xs, ohs, ds, v, lg, rate, nm, ns, 1.0, 0.004, 0.06, 1.2, 0.02, seed=3, step=1)


upd_ms = timed(upd, 5) / INNER
upd_bytes = 133 * int(dims_u.sum())   # dead slots are neither read nor written: 133 B per LIVE particle-step
gbs = upd_bytes / (upd_ms * 1e-3) / 1e9
del xs, ohs, v, lg, mask_u

# short sampler run
cfg.sampler_kwargs.dt = 0.02
sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
sampler = JumpSampler(structure=model.structure, **sk)
in_st = model.make_batch(torch.zeros(B, N, 3, device=dev), torch.zeros(B, N, S, device=dev), torch.full((B,), N, device=dev))
smp_ms = timed(lambda: sampler.sample(model.net, in_st, model.jump_diffusion_loss, jet_offset=0), 2, warm=1)
out = sampler.sample(model.net, in_st, model.jump_diffusion_loss, jet_offset=0)
n_eval = 50
print(json.dumps({"workload": f"C3 transepic: B={B}, N=128, S=8, bf16 stacks (2 x 2 blocks, 128 wide, 2 heads), trunk G=19",
                  "forward": {"ms_per_evaluation": fwd_ms, "jet_evals_per_s": B / (fwd_ms * 1e-3)},
                  "roofline_forward": {"kernel": "mmb_trans_forward (2 x absorb_head_tc_kernel + trunk)", "bound": "tensor", "achieved": tf,
                                       "peak": pk["bf16"], "unit": "TFLOP/s", "frac": tf / pk["bf16"],
                                       "algorithmic_flops_per_launch": flops, "peak_source": pk["src"]},
                  "roofline_update": {"kernel": "mmb::trans_sampler_onehot_kernel<8> + trans_sampler_update_kernel<8>", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                                      "unit": "GB/s", "frac": gbs / pk["hbm"], "ms_per_launch": upd_ms, "jets": BU, "l2": "inputs larger than L2, 10 launches per event pair",
                                      "algorithmic_bytes_per_launch": upd_bytes, "peak_source": pk["src"]},
                  "sampler": {"dt": 0.02, "evaluations": n_eval, "ms_per_run": smp_ms, "jets_per_s": B / (smp_ms * 1e-3),
                              "ms_per_step": smp_ms / n_eval, "mean_final_multiplicity": float(out.get_dims().float().mean())}}))
