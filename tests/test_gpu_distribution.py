"""GPU: the CUDA samplers (fp32, tcgen05 bf16, warp-MMA f16) against the distributions of the UNMODIFIED reference sampler
(tests/golden/mbm_distribution.npz; torch.poisson / torch's generator untouched — tests/golden/make_golden_distribution.py).
North star: "W1 on pT/eta/phi/jet mass and on flavor multiplicities within the reference's own seed-to-seed spread"."""
import numpy as np
import pytest
import torch

import distribution_lib as dl
from test_oracle_distribution import load
from multimodal_particles_b200 import HybridState
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.epic import as_u8
from multimodal_particles_b200.observables import jet_observables

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("precision", ["fp32", "bf16", "f16"])
def test_cuda_sampler_within_reference_seed_to_seed_spread(precision):
    z, cfg, model = load()
    model.to(DEV)
    B = int(z["jets"])
    batch = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(int(z["source_seed0"]) + 200))
    model.seed = 77
    state = HybridState(None, batch.source_continuous.clone(), batch.source_discrete.clone(), batch.source_mask.clone())
    out = model.simulate_dynamics(state, batch, precision=precision, jet_offset=0)
    stats = {"mean": z["stats_mean"].tolist(), "std": z["stats_std"].tolist()}
    _, fc, jets = jet_observables(out.continuous.to(DEV).contiguous(), as_u8(out.discrete.to(DEV)), as_u8(batch.source_mask.to(DEV)), stats)
    cand = dl.summarise(out.continuous.numpy(), out.discrete[..., 0].numpy(), batch.source_mask[..., 0].numpy(),
                        fc[..., 0].cpu().numpy(), jets.cpu().numpy(), int(z["nq"]))
    rows = dl.gate(z, cand, f"CUDA {precision}")
    assert len(rows) >= 17
    moved = (out.discrete != batch.source_discrete)[batch.source_mask.bool()].float().mean().item()
    assert 0.8 < moved < 0.95      # the reference runs moved 87.5 % of the live tokens
