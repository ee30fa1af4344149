"""Tiny invocations of every kernel family in one process (seconds on a GPU): a quick all-kernels smoke run, and the input for
`compute-sanitizer --tool memcheck python tools/sanitize_small.py` where the sanitizer is available (it is closed on this pool)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from multimodal_particles_b200 import HybridState, MultiModalBridgeMatching, _native
from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.config_classes.transdimensional_unconditional_config import TransdimensionalEpicConfig
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.observables import jet_observables
from multimodal_particles_b200.source import sample_source_state
from multimodal_particles_b200.states import AbsorbingBridgeState
from multimodal_particles_b200.transdimensional import JumpSampler, TransdimensionalJumpDiffusion
from multimodal_particles_b200.epic import as_u8

dev = "cuda:0"
torch.manual_seed(0)
cfg = MultimodalBridgeMatchingConfig(); cfg.bridge.num_timesteps = 6
model = MultiModalBridgeMatching(cfg).to(dev)
b = jetclass_like_databatch(9, generator=torch.Generator().manual_seed(1))
for prec in ("fp32", "bf16"):
    st = HybridState(None, b.source_continuous.clone(), b.source_discrete.clone(), b.source_mask.clone())
    out = model.simulate_dynamics(st, b, precision=prec)
    assert torch.isfinite(out.continuous).all()
acfg = AbsorbingConfig(); acfg.data.max_num_particles = 128; acfg.bridge.num_timesteps = 4
flow = AbsorbingFlow(acfg).to(dev)
st = AbsorbingBridgeState(None, b.source_continuous.clone(), b.source_discrete.clone(), b.source_mask.clone())
out = flow.simulate_dynamics(st, b)
assert torch.isfinite(out.continuous).all()
tcfg = TransdimensionalEpicConfig(); tcfg.sampler_kwargs.dt = 0.25
tm = TransdimensionalJumpDiffusion(tcfg).to(dev)
sk = {k: v for k, v in vars(tcfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
smp = JumpSampler(structure=tm.structure, **sk)
B, N, S = 5, 128, 8
st = smp.sample(tm.net, tm.make_batch(torch.zeros(B, N, 3, device=dev), torch.zeros(B, N, S, device=dev), torch.full((B,), N, device=dev)),
                tm.jump_diffusion_loss, jet_offset=0)
assert torch.isfinite(st.tuple_batch[0]).all()
x, k, m = sample_source_state(7, 128, target_multiplicity=np.arange(1, 100), compact=True)
jet_observables(x, k, m, {"mean": [1.0, 0, 0], "std": [0.3, 0.2, 0.2]})
s2 = model.sample_bridges(type("B", (), dict(source_continuous=b.source_continuous, source_discrete=b.source_discrete,
                                            target_continuous=b.source_continuous, target_discrete=b.source_discrete, target_mask=b.source_mask))())
torch.cuda.synchronize()
print("sanitize_small: ok")
