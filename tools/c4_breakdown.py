"""Where a solver step of the absorbing-flow loop (BASELINE config 4) goes, kernel family by kernel family, with CUDA events on
warm back-to-back launches (an ncu launch list times every launch cold and serialised, which inflates the short kernels ~2.5x).
Prints one JSON line.  Usage: python tools/c4_breakdown.py [B]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_particles_b200 import _native  # noqa: E402
from multimodal_particles_b200.absorbing_flows import AbsorbingFlow  # noqa: E402
from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig  # noqa: E402
from multimodal_particles_b200.databatch import jetclass_like_databatch  # noqa: E402
from multimodal_particles_b200.epic import as_u8  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = AbsorbingConfig()
cfg.data.max_num_particles, cfg.bridge.num_timesteps = 128, 100
torch.manual_seed(0)
flow = AbsorbingFlow(cfg).to(dev)
gen = flow.generator
b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(1234))
table = flow.step_table()
tb = gen.time_bias(table.t)
trunk, head = gen.native_trunk(dev), gen.native_head(dev)
x0, k0, m0 = b.source_continuous.to(dev).contiguous(), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


# whole loop
loop_ms = timed(lambda: _native.generate_absorbing(trunk, head, x0.clone(), k0.clone(), m0.clone(), table, tb, seed=1, jet_offset=0,
                                                   precision="bf16"), reps=3, warm=1)
# the state half way through (births so far), for the per-kernel timings
x, k, m = x0.clone(), k0.clone(), m0.clone()
_native.generate_absorbing(trunk, head, x, k, m, table, tb, seed=1, jet_offset=0, precision="bf16")
out = {"B": B, "loop_ms": loop_ms, "ms_per_step": loop_ms / table.n_steps, "live_start": float(m0.float().sum(1).mean()),
       "live_end": float(m.float().sum(1).mean())}
temb = table.temb[50:51].to(dev).contiguous()
for name, mask in (("start", m0), ("end", m)):
    xs = x0 if name == "start" else x
    ks = k0 if name == "start" else k
    v, lg, hid = trunk.forward(xs, ks, mask, temb, want_hidden=True, precision="bf16")
    tb1 = tb[50:51].to(dev)
    u = torch.rand(B, 128, device=dev)
    alog = head.forward(hid, mask, tb1)
    out[name] = {
        "trunk_forward_ms": timed(lambda: trunk.forward(xs, ks, mask, temb, want_hidden=True, precision="bf16")),
        "rate_head_ms": timed(lambda: head.forward(hid, mask, tb1)),
        "update_ms": timed(lambda: _native.bridge_update(xs.clone(), ks.clone(), mask.clone(), v, lg, u, 0.01, 5.0, 0.4, absorb_logit=alog,
                                                         u_absorb=u, sp=0.5, flags=_native.FLAG_MULTIMODAL | _native.FLAG_ABSORBING)),
        "three_clones_ms": timed(lambda: (xs.clone(), ks.clone(), mask.clone())),
    }
print(json.dumps(out))
