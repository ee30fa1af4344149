"""Cost of one jet in the warp-MMA generation engine as a function of its multiplicity: batches whose jets all have the same
number of live particles (prefix masks), f16 engine, C2 step count.  Prints us of one SM-warp-slot per jet, i.e. what a jet of
that size costs the chip, so that packing choices (jets per warp, warps per jet) can be judged from numbers.

    python tools/mma_cost_by_size.py [--jets 8192]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--jets", type=int, default=8192)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

from multimodal_particles_b200 import MultiModalBridgeMatching  # noqa: E402
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402

dev = torch.device("cuda:0")
cfg = MultimodalBridgeMatchingConfig()
cfg.bridge.num_timesteps = 100
torch.manual_seed(0)
model = MultiModalBridgeMatching(cfg).to(dev)
native = model.encoder.native_model(dev)
table = model.step_table()
B, N = args.jets, 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for m in (4, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 128):
    mask = (torch.arange(N, device=dev)[None, :] < m).to(torch.uint8).expand(B, N).contiguous()
    x0 = torch.randn(B, N, 3, device=dev) * mask[..., None]
    k0 = (torch.randint(0, 8, (B, N), device=dev, dtype=torch.uint8) * mask).contiguous()
    ms = []
    for i in range(args.reps + 2):
        x, k = x0.clone(), k0.clone()
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        native.generate(x, k, mask, table, seed=1, jet_offset=0, precision="f16")
        e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(s.elapsed_time(e))
    t = sorted(ms)[len(ms) // 2]
    out[m] = {"ms": t, "M_jets_per_s": B / t / 1e3, "ns_chip_per_jet": t * 1e6 / B}
    print(f"m={m:4d}: {t:.3f} ms  {B / t / 1e3:.2f} M jets/s  {t * 1e6 / B:.1f} ns of the chip per jet")
print(json.dumps(out))
