"""Bring-up check and timing of the warp-MMA generation engine (epic_mma.cu) against the fp32 path and the tcgen05 engine.

    python tools/mma_check.py [--lib path/to/libmmbridge.so] [--jets 4096] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--jets", type=int, default=4096)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--dense", action="store_true", help="also time a batch with all 128 particles live")
ap.add_argument("--only", default=None, help="time only this precision and skip the golden cases (profiling runs)")
args = ap.parse_args()
if args.lib:
    os.environ["MMB_LIB_PATH"] = os.path.abspath(args.lib)

import oracle_lib as ol  # noqa: E402
from multimodal_particles_b200 import HybridState, MultiModalBridgeMatching, _native  # noqa: E402
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402
from multimodal_particles_b200.databatch import jetclass_like_databatch  # noqa: E402
from multimodal_particles_b200.epic import as_u8  # noqa: E402

dev = torch.device("cuda:0")
out = {"lib": _native.LIB_PATH}
for case in (() if args.only else ("mbm_n128", "mbm_c1", "mbm_odd")):
    z, cfg, model = ol.load_mbm_golden(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    model.to(dev)
    st = lambda: HybridState(None, torch.from_numpy(z["x0"]), torch.from_numpy(z["k0"]).long(), torch.from_numpy(z["mask"]).long())
    u = torch.from_numpy(z["u_jump"])
    a = model.simulate_dynamics(st(), None, uniforms=u, precision="fp32")
    for prec in ("bf16", "f16"):
        try:
            b = model.simulate_dynamics(st(), None, uniforms=u, precision=prec)
        except Exception as exc:
            print(f"{case} {prec}: {exc}")
            continue
        live = torch.from_numpy(z["mask"]).bool()
        agree = (a.discrete == b.discrete)[live].float().mean().item()
        err = (a.continuous - b.continuous).abs()
        dead_ok = bool((b.discrete[~live] == 0).all() and (b.continuous[(~live).expand(-1, -1, 3)] == 0).all())
        print(f"{case} {prec}: token agreement(live)={agree:.4f} x max err={err.max():.5f} mean err={err.mean():.6f} "
              f"nan={int(torch.isnan(b.continuous).sum())} dead_zero={dead_ok}")
        out[f"{case}/{prec}"] = {"agree": agree, "x_max_err": err.max().item(), "x_mean_err": err.mean().item()}

# ---- timing at the C2 shape
cfg = MultimodalBridgeMatchingConfig()
cfg.bridge.num_timesteps = 100
torch.manual_seed(0)
model = MultiModalBridgeMatching(cfg).to(dev)
native = model.encoder.native_model(dev)
table = model.step_table()
B = args.jets
batch = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(1234))
x0, k0, m0 = batch.source_continuous.to(dev), as_u8(batch.source_discrete.to(dev)), as_u8(batch.source_mask.to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(prec, x0, k0, m0):
    res = None
    ms = []
    for i in range(args.reps + 2):
        x, k = x0.clone(), k0.clone()
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        native.generate(x, k, m0, table, seed=1, jet_offset=0, precision=prec)
        e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(s.elapsed_time(e))
        res = (x, k)
    return sum(ms) / len(ms), res


ref = None
for prec in ((args.only,) if args.only else ("fp32", "bf16", "f16")):
    try:
        ms, res = timed(prec, x0, k0, m0)
    except Exception as exc:
        print(f"C2 {prec}: {exc}")
        continue
    if prec == "fp32" or ref is None:
        ref = res
    live = m0.bool()
    agree = (res[1] == ref[1])[live].float().mean().item()
    err = (res[0] - ref[0]).abs().mean().item()
    print(f"C2 B={B} {prec}: {ms:.3f} ms -> {B / ms * 1e3 / 1e6:.3f} M jets/s; token agreement vs fp32 {agree:.4f}, mean |dx| {err:.5f}")
    out[f"C2/{prec}"] = {"ms": ms, "jets_per_s": B / ms * 1e3, "agree": agree, "mean_dx": err}
if args.dense:
    md = torch.ones_like(m0)
    xd = torch.randn(B, 128, 3, device=dev)
    kd = torch.randint(0, 8, (B, 128), device=dev, dtype=torch.uint8)
    for prec in ("bf16", "f16"):
        ms, _ = timed(prec, xd, kd, md)
        print(f"dense B={B} {prec}: {ms:.3f} ms -> {B / ms * 1e3 / 1e6:.3f} M jets/s")
        out[f"dense/{prec}"] = {"ms": ms, "jets_per_s": B / ms * 1e3}
print(json.dumps(out))
