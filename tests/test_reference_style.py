"""The reference's own generative tests (/root/reference/tests/test_generative/test_{multimodal,absorbing,transdimensional}.py),
re-stated against this package: same calls in the same order, so a reference user can read them side by side.
Config round trips run on the CPU; everything that evaluates a network or a bridge needs the GPU library."""
import os

import pytest
import torch

from multimodal_particles_b200 import AbsorbingBridgeState, HybridState, MultiModalBridgeMatching
from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
from multimodal_particles_b200.bridges import AbsorbingBridge
from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.config_classes.transdimensional_unconditional_config import TransdimensionalEpicConfig
from multimodal_particles_b200.databatch import random_databatch
from multimodal_particles_b200.transdimensional import TransdimensionalJumpDiffusion

DEV = "cuda:0"


def on_device(batch):
    return type(batch)(*[t.to(DEV) for t in batch])


# ---- test_multimodal.py / test_absorbing.py / test_transdimensional.py :: test_config(s) -------------------------------
@pytest.mark.parametrize("cls", [MultimodalBridgeMatchingConfig, AbsorbingConfig, TransdimensionalEpicConfig])
def test_config(cls, tmp_path):
    path = os.path.join(tmp_path, "config.yaml")
    config = cls()
    config.to_yaml(path)
    config_read = cls.from_yaml(path)
    assert config_read is not None and config_read.data.max_num_particles == config.data.max_num_particles


# ---- test_multimodal.py :: test_model ------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_model():
    model_config = MultimodalBridgeMatchingConfig()
    random_batch = on_device(random_databatch(model_config))
    model = MultiModalBridgeMatching(model_config).to(DEV)
    state = model.sample_bridges(random_batch)
    head_output = model(state, random_batch)
    B, N = model_config.data.batch_size, model_config.data.max_num_particles
    assert state.time.shape == (B, 1, 1) and state.continuous.shape == (B, N, 3) and state.discrete.shape == (B, N, 1)
    assert state.absorbing.shape == (B, N, 1)
    assert head_output.continuous.shape == (B, N, 3) and head_output.discrete.shape == (B, N, model_config.data.vocab_size_features)


# ---- test_absorbing.py :: test_bridge -------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_bridge():
    config = AbsorbingConfig()
    model = AbsorbingFlow(config).to(DEV)
    random_batch = on_device(random_databatch(config))
    absorbing_bridge = AbsorbingBridge(config)
    # all equal to target at time 1
    t = torch.ones(random_batch.target_continuous.shape[0], device=DEV).type_as(random_batch.target_continuous)
    time = model.reshape_time(t, random_batch.target_continuous)
    mask_t = absorbing_bridge.sample(time, random_batch.target_mask)
    assert (mask_t == random_batch.target_mask).all()
    # all existing at time 0
    t = torch.zeros(random_batch.target_continuous.shape[0], device=DEV).type_as(random_batch.target_continuous)
    time = model.reshape_time(t, random_batch.target_continuous)
    mask_t = absorbing_bridge.sample(time, random_batch.target_mask)
    assert (mask_t == 1).all()
    # sample full state
    state = model.sample_bridges(random_batch)
    assert state is not None and state.mask_t.shape == random_batch.target_mask.shape


# ---- test_absorbing.py :: test_absorbing_head / test_absorbing_dynamics -----------------------------------------------------
@pytest.mark.gpu
def test_absorbing_dynamics():
    config = AbsorbingConfig()
    config.bridge.num_timesteps = 20         # the reference runs its default 1000 steps on the CPU; the loop is the same
    model = AbsorbingFlow(config).to(DEV)
    random_batch = on_device(random_databatch(config))
    # start in the source
    initial_state = AbsorbingBridgeState(None, random_batch.source_continuous, random_batch.source_discrete, random_batch.source_mask)
    # one absorbing step
    initial_state.time = torch.full((initial_state.continuous.size(0), 1), 0.01, device=DEV)
    heads = model.forward(initial_state, random_batch)
    B, N = config.data.batch_size, config.data.max_num_particles
    assert heads.absorbing.shape == (B, N, 1) and heads.continuous.shape == (B, N, 3)
    next_state = model.bridge_absorbing.solver_step(initial_state, heads, 0.01)
    assert next_state is initial_state and (next_state.mask_t >= random_batch.source_mask).all()     # in place; particles are only born
    # simulate dynamics
    final_state = model.simulate_dynamics(initial_state, random_batch)
    assert final_state.continuous.device.type == "cpu" and final_state.mask_t.shape == (B, N, 1)


# ---- test_transdimensional.py :: test_model -----------------------------------------------------------------------------------
@pytest.mark.gpu
def test_transdimensional_model():
    config = TransdimensionalEpicConfig()
    config.data.return_type = "list"
    model = TransdimensionalJumpDiffusion(config).to(DEV)
    B, N, S = 12, config.data.max_num_particles, config.data.vocab_size_features
    g = torch.Generator().manual_seed(0)
    dims = torch.randint(1, N + 1, (B,), generator=g)
    m = (torch.arange(N)[None] < dims[:, None]).float().unsqueeze(-1)
    st_batch = model.make_batch((torch.randn(B, N, 3, generator=g) * m).to(DEV), (torch.randn(B, N, S, generator=g) * m).to(DEV), dims.to(DEV))
    ts = (config.loss_kwargs.min_t + (1 - config.loss_kwargs.min_t) * torch.rand(B, generator=g)).to(DEV)
    D_xt, rate_xt, dummy_mean_std, x0_dim_logits, _ = model.net(st_batch, ts=ts, forward_rate=model.forward_rate, predict="eps",
                                                                 nearest_atom=torch.zeros((B,), device=DEV).long())
    assert rate_xt.shape == (B, 1) and D_xt.shape == (B, N * (3 + S)) and x0_dim_logits.shape == (B, N)
    mean, std = model.noise_schedule.get_p0t_stats(st_batch, ts)
    assert mean.shape == D_xt.shape and std.shape == D_xt.shape
