"""Golden fixture of the absorbing flow, produced by the UNMODIFIED reference (container only):

    python tests/golden/make_golden_absorbing.py

AbsorbingFlow.simulate_dynamics (mp/models/generative/absorbing/absorbing_flows.py:255-275) with
``torch.poisson`` / ``torch.bernoulli`` replaced by the uniform-driven equivalents of
make_golden.py.  Records, per selected step, the state fed to the generator, the three heads and the
per-block time biases, plus the mask/token trajectories and the final state.
"""
import json
import os
import sys
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shim)

from multimodal_particles.config_classes.absorbing_flows_config import AbsorbingConfig  # noqa: E402
from multimodal_particles.models.generative.absorbing.absorbing_flows import AbsorbingFlow  # noqa: E402
from multimodal_particles.models.generative.absorbing.states import AbsorbingBridgeState  # noqa: E402
from multimodal_particles.models.architectures.gsdm import get_timestep_embedding, nonlinearity  # noqa: E402


def main():
    cfg = AbsorbingConfig()
    cfg.data.max_num_particles, cfg.data.batch_size = 24, 4
    cfg.bridge.num_timesteps = 13
    torch.manual_seed(201)
    model = AbsorbingFlow(cfg)
    with torch.no_grad():  # make births, moves and rates non-trivial at random init
        model.generator.discrete_head_mlp[2].weight.mul_(4.0)
        model.generator.post_rate_proj.weight.mul_(6.0)
        model.generator.epic.epic.output_layer.weight_g.mul_(3.0)
    B, N, S = 4, 24, cfg.data.vocab_size_features
    steps = cfg.bridge.num_timesteps - 1
    g = torch.Generator().manual_seed(202)
    mask = mg.prefix_masks([1, 7, 12, 20], N)
    x0 = torch.randn(B, N, 3, generator=g) * mask
    k0 = torch.randint(0, S, (B, N, 1), generator=g) * mask
    uj, ua = torch.rand(steps, B, N, generator=g), torch.rand(steps, B, N, generator=g)
    ua = ua * 0.02  # births are rare per step (p ~ dt * SP * sigmoid): bias the draws so masks do grow

    rec = {"t": [], "snap": {}, "mask_traj": [], "k_traj": []}
    snap_steps = {0, 5, 11}
    orig_forward = model.forward

    def forward(state, batch):
        i = len(rec["t"])
        heads = orig_forward(state, batch)
        rec["t"].append(state.time[0, 0].item())
        if i in snap_steps:
            ts = state.time.squeeze()
            temb = model.generator.temb_net(get_timestep_embedding(ts * 1000, model.generator.temb_dim))
            tb = torch.stack([blk.temb_proj(nonlinearity(temb)[:, :, None])[:, :, 0] for blk in model.generator.res_blocks], 1)
            rec["snap"][i] = dict(x=state.continuous.clone(), k=state.discrete.clone(), mask=state.mask_t.clone(),
                                  v=heads.continuous.detach().clone(), logits=heads.discrete.detach().clone(),
                                  a=heads.absorbing.detach().clone(), tbias=tb.detach().clone())
        return heads

    model.forward = forward
    orig_jump = model.bridge_discrete.solver_step

    def jump(state, heads, dt, multimodal=True):
        out = orig_jump(state, heads, dt, multimodal=multimodal)
        rec["k_traj"].append(out.discrete.squeeze(-1).clone())
        rec["mask_traj"].append(out.mask_t.squeeze(-1).clone())
        return out

    model.bridge_discrete.solver_step = jump
    state = AbsorbingBridgeState(None, x0.clone(), k0.clone(), mask.clone())
    with mg.InjectedNoise(u_jump=uj, u_absorb=ua), torch.no_grad():
        final = model.simulate_dynamics(state, (x0,))

    t_all = torch.tensor(rec["t"], dtype=torch.float32)
    temb_all = model.generator.temb_net(get_timestep_embedding(t_all * 1000, model.generator.temb_dim))
    tb_all = torch.stack([blk.temb_proj(nonlinearity(temb_all)[:, :, None])[:, :, 0] for blk in model.generator.res_blocks], 1)
    out = dict(config=json.dumps(asdict(cfg)), x0=x0.numpy(), k0=k0.numpy().astype(np.uint8), mask0=mask.numpy().astype(np.uint8),
               u_jump=uj.numpy(), u_absorb=ua.numpy(), t=t_all.numpy(), tbias=tb_all.detach().numpy(),
               sp=model.bridge_absorbing.survival_probability(t_all).numpy(),
               k_traj=torch.stack(rec["k_traj"]).numpy().astype(np.uint8),
               mask_traj=torch.stack(rec["mask_traj"]).numpy().astype(np.uint8),
               x_final=final.continuous.numpy(), k_final=final.discrete.numpy().astype(np.uint8),
               mask_final=final.mask_t.numpy().astype(np.uint8), snap_steps=np.array(sorted(rec["snap"]), dtype=np.int32))
    for i, s in rec["snap"].items():
        for key, val in s.items():
            arr = val.numpy()
            out[f"snap{i}/{key}"] = arr.astype(np.uint8) if arr.dtype == np.int64 else arr
    out.update(mg.np_state_dict(model))
    path = os.path.join(HERE, "absorbing.npz")
    np.savez_compressed(path, **out)
    born = int(final.mask_t.sum() - mask.sum())
    print(f"absorbing: {os.path.getsize(path) / 1024:.0f} KiB, particles born: {born}, "
          f"token moves: {(torch.stack(rec['k_traj'])[1:] != torch.stack(rec['k_traj'])[:-1]).sum().item()}")


if __name__ == "__main__":
    main()
