"""Bring-up check of the tcgen05 absorbing-rate head against the fp32 oracle (prints error statistics)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from multimodal_particles_b200.states import AbsorbingBridgeState  # noqa: E402

dev = torch.device("cuda:0")
z, cfg, model = ol.load_absorbing_golden(os.path.join(ROOT, "tests", "golden", "absorbing.npz"))
g = model.generator
blob = g.pack_head_weights().numpy()
trunk = ol.absorbing_trunk(model)
tab = model.step_table()
model.to(dev)
for i in z["snap_steps"]:
    s = lambda name: z[f"snap{i}/{name}"]
    v, logits, hidden = ol.epic_forward(*trunk, s("x"), s("k")[..., 0], s("mask")[..., 0], tab.temb[i].numpy()[None], want_hidden=True)
    want = ol.absorb_head(blob, 16, 128, 2, 2, hidden, s("mask")[..., 0], s("tbias")[:1])
    head = g.native_head(dev)
    got = head.forward(torch.from_numpy(hidden).to(dev), torch.from_numpy(s("mask")[..., 0]).to(dev), torch.from_numpy(s("tbias")[:1]).to(dev))
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    print(f"step {i}: rate logit max|ref|={np.abs(want).max():.4f} max err={np.abs(got - want).max():.5f} "
          f"mean err={np.abs(got - want).mean():.6f} nan={int(np.isnan(got).sum())} vs reference {np.abs(got - s('a')[..., 0]).max():.5f}")
st = AbsorbingBridgeState(None, torch.from_numpy(z["x0"]), torch.from_numpy(z["k0"]).long(), torch.from_numpy(z["mask0"]).long())
for prec in ("fp32", "bf16"):
    t0 = time.perf_counter()
    out = model.simulate_dynamics(AbsorbingBridgeState(None, st.continuous.clone(), st.discrete.clone(), st.mask_t.clone()), None,
                                  uniforms_jump=torch.from_numpy(z["u_jump"]), uniforms_absorb=torch.from_numpy(z["u_absorb"]), precision=prec)
    dt = time.perf_counter() - t0
    m_ok = (out.mask_t.numpy() == z["mask_final"]).mean()
    k_ok = (out.discrete.numpy() == z["k_final"]).mean()
    print(f"generate[{prec}]: mask agreement {m_ok:.4f}, token agreement {k_ok:.4f}, x max err {np.abs(out.continuous.numpy() - z['x_final']).max():.4f} ({dt:.3f}s)")
# throughput probe: 512 jets x 128 particles, one head evaluation
B, N = 1184, 128
hid = torch.randn(B, N, 16, device=dev)
msk = (torch.rand(B, N, device=dev) < 0.4).to(torch.uint8)
tb = torch.randn(1, 2, 128, device=dev) * 0.1
head = g.native_head(dev)
for _ in range(2):
    head.forward(hid, msk, tb)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    head.forward(hid, msk, tb)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print(f"head: {B} jets in {ms:.3f} ms -> {B / ms * 1e3:.0f} jet-evals/s, {72.0e6 * B / ms / 1e9:.1f} TFLOP/s")
