"""Throughput of the absorbing flow (BASELINE config 4, absorbing part): AbsorbingFlow.simulate_dynamics
on synthetic JetClass-shaped jets, 99 solver steps, plus the rate-head kernel alone against the bf16
tensor roofline (72.0 MFLOP per jet-step).  Prints one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from multimodal_particles_b200.absorbing_flows import AbsorbingFlow  # noqa: E402
from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig  # noqa: E402
from multimodal_particles_b200.databatch import jetclass_like_databatch  # noqa: E402
from multimodal_particles_b200.epic import as_u8  # noqa: E402
from multimodal_particles_b200.states import AbsorbingBridgeState  # noqa: E402
from multimodal_particles_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = AbsorbingConfig()
cfg.data.max_num_particles, cfg.bridge.num_timesteps = 128, 100
torch.manual_seed(0)
model = AbsorbingFlow(cfg).to(dev)
gen = model.generator
b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(1234))
table = model.step_table()
tb = gen.time_bias(table.t)
trunk, head = gen.native_trunk(dev), gen.native_head(dev)
x0, k0, m0 = b.source_continuous.to(dev).contiguous(), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
times = []
for i in range(5):
    x, k, m = x0.clone(), k0.clone(), m0.clone()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _native.generate_absorbing(trunk, head, x, k, m, table, tb, seed=1, jet_offset=0, precision="bf16")
    e.record()
    torch.cuda.synchronize()
    if i >= 2:
        times.append(s.elapsed_time(e))
ms = sum(times) / len(times)
# e2e through the public API with pinned host tensors
st = lambda: AbsorbingBridgeState(None, b.source_continuous.clone().pin_memory(), b.source_discrete.clone().pin_memory(),
                                  b.source_mask.clone().pin_memory())
model.simulate_dynamics(st(), b)
t0 = time.perf_counter()
out = model.simulate_dynamics(st(), b)
torch.cuda.synchronize()
e2e_s = time.perf_counter() - t0
# rate head alone
hid = torch.randn(B, 128, 16, device=dev)
tb1 = tb[:1].to(dev)
for _ in range(3):
    head.forward(hid, m0, tb1)
ht = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); head.forward(hid, m0, tb1); e.record(); torch.cuda.synchronize()
    ht.append(s.elapsed_time(e))
hms = sum(ht) / len(ht)
pk = bench.peaks()
tf = 72.0e6 * B / (hms * 1e-3) / 1e12
print(json.dumps({"workload": f"C4 absorbing flow: B={B}, N=128, 99 steps, EPiC trunk + 128-wide 2-block transformer rate head",
                  "value": B / (ms * 1e-3), "unit": "jets/s", "ms_per_generation": ms,
                  "e2e": {"value": B / e2e_s, "unit": "jets/s"},
                  "births": int(out.mask_t.sum() - b.source_mask.sum()),
                  "roofline_head": {"kernel": "mmb::absorb_head_tc_kernel", "bound": "tensor", "achieved": tf, "peak": pk["bf16"],
                                    "unit": "TFLOP/s", "frac": tf / pk["bf16"], "ms_per_launch": hms,
                                    "algorithmic_flops_per_launch": 72.0e6 * B, "peak_source": pk["src"]}}))
