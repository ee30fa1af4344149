"""Bring-up check of the tcgen05 path against the fp32 path (prints error statistics)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from multimodal_particles_b200 import HybridState  # noqa: E402

dev = torch.device("cuda:0")
for case in ("mbm_n128", "mbm_c1"):
    z, cfg, model = ol.load_mbm_golden(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    model.to(dev)
    native = model.encoder.native_model(dev)
    i = int(z["snap_steps"][1])
    x, k, mask = (torch.from_numpy(z[f"snap{i}/x"]).to(dev), torch.from_numpy(z[f"snap{i}/k"][..., 0]).to(dev),
                  torch.from_numpy(z["mask"][..., 0]).to(dev))
    temb = torch.from_numpy(z["temb"][i][None]).to(dev)
    ref = native.forward(x, k, mask, temb, want_hidden=True, precision="fp32")
    got = native.forward(x, k, mask, temb, want_hidden=True, precision="bf16")
    torch.cuda.synchronize()
    for name, a, b in zip(("v", "logits", "hidden"), ref, got):
        err = (a - b).abs()
        print(f"{case} step {i} {name}: max|ref|={a.abs().max():.4f} max err={err.max():.5f} mean err={err.mean():.6f} "
              f"nan={int(torch.isnan(b).sum())}")
    # whole generation with injected uniforms
    st = lambda: HybridState(None, torch.from_numpy(z["x0"]), torch.from_numpy(z["k0"]).long(), torch.from_numpy(z["mask"]).long())
    u = torch.from_numpy(z["u_jump"])
    a = model.simulate_dynamics(st(), None, uniforms=u, precision="fp32")
    t0 = time.perf_counter()
    b = model.simulate_dynamics(st(), None, uniforms=u, precision="bf16")
    print(f"{case} generate: token agreement={(a.discrete == b.discrete).float().mean():.4f} "
          f"x max err={(a.continuous - b.continuous).abs().max():.4f} mean err={(a.continuous - b.continuous).abs().mean():.5f} "
          f"({time.perf_counter() - t0:.3f}s)")
