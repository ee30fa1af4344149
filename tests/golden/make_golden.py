"""Generate the committed golden fixtures by executing the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference consumes torch's global generator through ``torch.poisson`` / ``torch.bernoulli``
with implementation-defined streams (SURVEY.md §7 "RNG"), so stream-level reproduction is
impossible by construction.  The fixtures therefore come from the reference run with those two
functions replaced by deterministic functions of PRE-DRAWN uniforms that realise the same
distributions:

* ``torch.poisson(lam)`` -> one-hot jump counts chosen by one uniform per particle with
  ``P(select s) = lam_s exp(-sum lam)`` (SURVEY.md §A.4 Form B, self slot retained).  Everything
  downstream — gate ``sum J <= 1``, ``k += sum J (s-k)``, clamp, mask — is the reference's own code;
* ``torch.bernoulli(p)`` -> ``(u < p)``.

Each .npz holds the config, the state dict, the inputs, the uniforms, the reference's per-step
times / heads at selected steps, the token trajectory and the final state.
"""
import json
import os
import sys
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from multimodal_particles.config_classes.multimodal_bridge_matching_config import (  # noqa: E402
    MultimodalBridgeMatchingConfig)
from multimodal_particles.models.generative.multimodal_bridge_matching import (  # noqa: E402
    HybridState, MultiModalBridgeMatching)
from multimodal_particles.models.generative import bridges as ref_bridges  # noqa: E402


class InjectedNoise:
    """Context manager swapping torch.poisson / torch.bernoulli for uniform-driven versions."""

    def __init__(self, u_jump=None, u_absorb=None):
        self.u_jump, self.u_absorb = u_jump, u_absorb
        self.i_jump = self.i_absorb = 0

    def poisson(self, lam):
        u = self.u_jump[self.i_jump]
        self.i_jump += 1
        S = lam.shape[-1]
        total = lam[..., 0]
        for s in range(1, S):
            total = total + lam[..., s]
        e = torch.exp(-total)
        c = torch.zeros_like(total)
        chosen = torch.full(total.shape, -1, dtype=torch.long)
        for s in range(S):
            c = c + lam[..., s] * e
            hit = (u < c) & (chosen < 0)
            chosen[hit] = s
        jumps = torch.zeros_like(lam)
        sel = chosen >= 0
        jumps[sel] = torch.nn.functional.one_hot(chosen[sel], S).to(lam.dtype)
        return jumps

    def bernoulli(self, p):
        u = self.u_absorb[self.i_absorb]
        self.i_absorb += 1
        return (u < p).to(p.dtype)

    def __enter__(self):
        self._poisson, self._bernoulli = torch.poisson, torch.bernoulli
        torch.poisson, torch.bernoulli = self.poisson, self.bernoulli
        return self

    def __exit__(self, *exc):
        torch.poisson, torch.bernoulli = self._poisson, self._bernoulli


def np_state_dict(module):
    return {"sd/" + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def prefix_masks(mults, n):
    return (torch.arange(n)[None, :] < torch.tensor(mults)[:, None]).long().unsqueeze(-1)


def mbm_case(name, cfg, x0, k0, mask, seed, snap_steps, batch=None, extra=None):
    """``batch``: what the reference's forward reads context features from (mbm.py:143-144); ``extra``: arrays to store too."""
    torch.manual_seed(seed)
    model = MultiModalBridgeMatching(cfg)
    # random init leaves logits nearly flat; scale the head so tokens actually compete
    with torch.no_grad():
        if cfg.encoder.add_discrete_head:
            model.encoder.fc_layer[2].weight.mul_(6.0)
        model.encoder.epic.epic.output_layer.weight_g.mul_(3.0)
    B, N, _ = x0.shape
    steps = cfg.bridge.num_timesteps - 1
    g = torch.Generator().manual_seed(seed + 1)
    u = torch.rand(steps, B, N, generator=g)

    rec = {"t": [], "k_traj": [], "snap": {}}
    orig_forward = model.forward

    def forward(state, batch):
        i = len(rec["t"])
        heads = orig_forward(state, batch)
        rec["t"].append(state.time[0, 0].item())
        if i in snap_steps:
            rec["snap"][i] = dict(x=state.continuous.clone(), k=state.discrete.clone(),
                                  v=heads.continuous.detach().clone(), logits=heads.discrete.detach().clone())
        return heads

    model.forward = forward
    orig_jump = model.bridge_discrete.solver_step

    def jump(state, heads, dt):
        out = orig_jump(state, heads, dt)
        rec["k_traj"].append(out.discrete.squeeze(-1).clone())
        return out

    model.bridge_discrete.solver_step = jump
    batch = (x0,) if batch is None else batch
    state = HybridState(None, x0.clone(), k0.clone(), mask.clone())
    with InjectedNoise(u_jump=u), torch.no_grad():
        final = model.simulate_dynamics(state, batch)

    # the reference's own per-step coefficients, for the step-table test
    t_all = torch.tensor(rec["t"], dtype=torch.float32)
    S = cfg.data.vocab_size_features
    wt = torch.exp(-S * cfg.bridge.gamma * (1.0 - t_all))
    temb = model.encoder.epic.embedding.embedding_time(t_all)
    grid = torch.linspace(0.0, 1.0 - cfg.bridge.time_eps, cfg.bridge.num_timesteps)
    dt = (grid[-1] - grid[0]) / (len(grid) - 1)

    out = dict(config=json.dumps(asdict(cfg)), x0=x0.numpy(), k0=k0.numpy().astype(np.uint8),
               mask=mask.numpy().astype(np.uint8), u_jump=u.numpy(),
               t=t_all.numpy(), temb=temb.numpy(), bc=((wt * S) / (1.0 - wt)).numpy(), cc=wt.numpy(),
               dt=np.float32(dt.item()),
               k_traj=torch.stack(rec["k_traj"]).numpy().astype(np.uint8),
               x_final=final.continuous.numpy(), k_final=final.discrete.numpy().astype(np.uint8),
               snap_steps=np.array(sorted(rec["snap"]), dtype=np.int32))
    for i, s in rec["snap"].items():
        out[f"snap{i}/x"] = s["x"].numpy()
        out[f"snap{i}/k"] = s["k"].numpy().astype(np.uint8)
        out[f"snap{i}/v"] = s["v"].numpy()
        out[f"snap{i}/logits"] = s["logits"].numpy()
    out.update(np_state_dict(model))
    out.update(extra or {})
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    moved = (torch.stack(rec["k_traj"])[1:] != torch.stack(rec["k_traj"])[:-1]).sum().item()
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, token moves over trajectory: {moved}")


def update_case(name, seed):
    """The two/three solver_steps alone on wide-ranging heads (bridges.py:38-45,179-201,260-286)."""
    from multimodal_particles.config_classes.absorbing_flows_config import AbsorbingConfig
    from multimodal_particles.models.generative.absorbing.states import AbsorbingBridgeState, OutputHeads
    cfg = AbsorbingConfig()
    cfg.bridge.num_timesteps = 100
    S, B, N, Dc = cfg.data.vocab_size_features, 6, 64, 3
    g = torch.Generator().manual_seed(seed)
    cont, tele, absb = (ref_bridges.LinearUniformBridge(cfg), ref_bridges.TelegraphBridge(cfg),
                        ref_bridges.AbsorbingBridge(cfg))
    grid = torch.linspace(0.0, 1.0 - cfg.bridge.time_eps, cfg.bridge.num_timesteps)
    dt = (grid[-1] - grid[0]) / (len(grid) - 1)
    out = dict(dt=np.float32(dt.item()), gamma=np.float32(cfg.bridge.gamma),
               gamma_absorb=np.float32(cfg.bridge.gamma_absorb))
    steps = [1, 40, 90, 97, 98, 99]
    out["steps"] = np.array(steps, dtype=np.int32)
    for i in steps:
        t = grid[i]
        x = torch.randn(B, N, Dc, generator=g) * 2
        k = torch.randint(0, S, (B, N, 1), generator=g)
        mask = torch.randint(0, 2, (B, N, 1), generator=g)
        v = torch.randn(B, N, Dc, generator=g) * 3
        logits = torch.randn(B, N, S, generator=g) * (4.0 if i % 2 else 1.0)
        a = torch.randn(B, N, 1, generator=g) * 3
        uj = torch.rand(1, B, N, generator=g)
        ua = torch.rand(1, B, N, generator=g)
        # bias some uniforms low so late-step (tiny) move probabilities are still exercised
        uj[0, :, ::3] *= 0.02
        time = torch.full((B, 1), t.item())
        wt = torch.exp(-S * cfg.bridge.gamma * (1.0 - time.squeeze()))
        for mode in ("mbm", "abs"):
            if mode == "mbm":
                from multimodal_particles.models.generative.multimodal_bridge_matching import MultiHeadOutput
                st = HybridState(time, x.clone(), k.clone(), mask.clone())
                heads = MultiHeadOutput(v, logits, mask)
                with InjectedNoise(u_jump=uj):
                    st = cont.solver_step(st, heads, dt)
                    st = tele.solver_step(st, heads, dt)
                res = dict(x=st.continuous, k=st.discrete, mask=mask)
            else:
                st = AbsorbingBridgeState(time, x.clone(), k.clone(), mask.clone())
                heads = OutputHeads(v, logits, a)
                with InjectedNoise(u_jump=uj, u_absorb=ua):
                    st = absb.solver_step(st, heads, dt)
                    st = cont.solver_step(st, heads, dt, multimodal=False)
                    st = tele.solver_step(st, heads, dt, multimodal=False)
                res = dict(x=st.continuous, k=st.discrete, mask=st.mask_t)
            for key, val in res.items():
                arr = val.numpy()
                out[f"s{i}/{mode}/{key}"] = arr.astype(np.uint8) if arr.dtype == np.int64 else arr
        out[f"s{i}/in/x"], out[f"s{i}/in/k"], out[f"s{i}/in/mask"] = (x.numpy(), k.numpy().astype(np.uint8),
                                                                      mask.numpy().astype(np.uint8))
        out[f"s{i}/in/v"], out[f"s{i}/in/logits"], out[f"s{i}/in/a"] = v.numpy(), logits.numpy(), a.numpy()
        out[f"s{i}/in/uj"], out[f"s{i}/in/ua"] = uj[0].numpy(), ua[0].numpy()
        out[f"s{i}/t"] = np.float32(t.item())
        out[f"s{i}/bc"] = ((wt * S) / (1.0 - wt))[0].numpy()
        out[f"s{i}/cc"] = wt[0].numpy()
        out[f"s{i}/sp"] = absb.survival_probability(time)[0, 0].numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    ref_cfg = "/root/reference/tests/resources/configs_files/config-mbm-test.yaml"

    # C1: toy shape of tests/test_generative (N=30, S=4), Bernoulli masks as random_databatch
    cfg = MultimodalBridgeMatchingConfig.from_yaml(ref_cfg)
    cfg.data.max_num_particles, cfg.data.vocab_size_features, cfg.data.batch_size = 30, 4, 6
    g = torch.Generator().manual_seed(11)
    B, N, S = 6, 30, 4
    mask = torch.randint(0, 2, (B, N, 1), generator=g)
    mask[:, 0] = 1
    mbm_case("mbm_c1", cfg, torch.rand(B, N, 3, generator=g), torch.randint(0, S, (B, N, 1), generator=g), mask,
             seed=101, snap_steps={0, 1, 50, 97, 98})

    # config-mbm-test shape (N=128, S=8), JetClass-like prefix masks incl. 1 and 128 particles
    cfg = MultimodalBridgeMatchingConfig.from_yaml(ref_cfg)
    B, N, S = 4, 128, 8
    g = torch.Generator().manual_seed(12)
    mask = prefix_masks([128, 1, 45, 77], N)
    x0 = torch.randn(B, N, 3, generator=g) * mask
    k0 = torch.randint(0, S, (B, N, 1), generator=g) * mask
    mbm_case("mbm_n128", cfg, x0, k0, mask, seed=102, snap_steps={0, 49, 98})

    # odd widths (transepic uses G=19), no skip connection, no discrete head, 3 blocks
    cfg = MultimodalBridgeMatchingConfig.from_yaml(ref_cfg)
    e, d = cfg.encoder, cfg.data
    e.dim_hidden_glob, e.dim_hidden_local, e.dim_emb_time, e.num_blocks = 19, 24, 14, 3
    e.dim_emb_features_continuous, e.dim_emb_features_discrete = 12, 10
    e.skip_connection, e.add_discrete_head = False, False
    d.max_num_particles, d.vocab_size_features, d.dim_features_continuous = 37, 5, 2
    cfg.bridge.num_timesteps = 12
    B, N, S = 3, 37, 5
    g = torch.Generator().manual_seed(13)
    mask = prefix_masks([37, 20, 5], N)
    mbm_case("mbm_odd", cfg, torch.randn(B, N, 2, generator=g) * mask,
             torch.randint(0, S, (B, N, 1), generator=g) * mask, mask, seed=103, snap_steps={0, 5, 10})

    update_case("bridge_update", seed=21)


if __name__ == "__main__":
    main()
