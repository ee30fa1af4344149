"""Forward half of a training / validation step: oracle against the reference fixture (CPU), kernels against
the oracle and the fixture (GPU).  tests/golden/make_golden_train.py produced the fixture with the reference's own
sample_bridges / loss_continuous / loss_discrete / AbsorbingBridge.sample and injected draws."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_forward.npz")


def test_oracle_matches_reference_bridges_and_losses():
    z = np.load(GOLD)
    xt, kt = ol.sample_bridges(z["x0"], z["x1"], z["k0"][..., 0], z["k1"][..., 0], z["t"], float(z["sigma"]), float(z["gamma"]), 8,
                               z["z"], z["u"])
    np.testing.assert_allclose(xt, z["xt"], rtol=0, atol=1e-6)
    assert np.array_equal(kt, z["kt"][..., 0])
    assert (kt != z["k0"][..., 0]).any() and (kt != z["k1"][..., 0]).any()          # the bridge really interpolates
    losses = ol.bridge_losses(z["v"], z["logits"], z["x0"], z["x1"], z["k1"][..., 0], z["mask"][..., 0])
    np.testing.assert_allclose(losses[:2], [z["loss_continuous"], z["loss_discrete"]], rtol=2e-6)
    assert losses[2] == z["mask"].sum()
    mt = ol.absorbing_sample(z["sp"], z["mask"][..., 0], z["u_absorb"][..., 0])
    assert np.array_equal(mt, z["mask_t"][..., 0]) and (mt >= z["mask"][..., 0]).all() and mt.sum() > z["mask"].sum()


@pytest.mark.gpu
def test_kernels_match_oracle_and_reference():
    from multimodal_particles_b200 import _native
    z = np.load(GOLD)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
    xt, kt = _native.sample_bridges(dev(z["x0"]), dev(z["x1"]), dev(z["k0"][..., 0]), dev(z["k1"][..., 0]), dev(z["t"]), float(z["sigma"]),
                                    float(z["gamma"]), 8, dev(z["z"]), dev(z["u"]))
    wxt, wkt = ol.sample_bridges(z["x0"], z["x1"], z["k0"][..., 0], z["k1"][..., 0], z["t"], float(z["sigma"]), float(z["gamma"]), 8,
                                 z["z"], z["u"])
    assert np.array_equal(xt.cpu().numpy(), wxt) and np.array_equal(kt.cpu().numpy(), wkt)      # bit-exact against the oracle
    assert np.array_equal(kt.cpu().numpy(), z["kt"][..., 0])
    out = _native.bridge_losses(dev(z["v"]), dev(z["logits"]), dev(z["x0"]), dev(z["x1"]), dev(z["k1"][..., 0]), dev(z["mask"][..., 0]))
    np.testing.assert_allclose(out.cpu().numpy()[:2], [z["loss_continuous"], z["loss_discrete"]], rtol=1e-5)
    mt = _native.absorbing_sample(dev(z["sp"]), dev(z["mask"][..., 0]), dev(z["u_absorb"][..., 0]))
    assert np.array_equal(mt.cpu().numpy(), z["mask_t"][..., 0])


@pytest.mark.gpu
def test_mirror_api_and_full_size_properties():
    """reference-shaped calls (sample_bridges / loss_* / validation_step / AbsorbingBridge.sample) at the C2 batch size"""
    from multimodal_particles_b200 import MultiModalBridgeMatching
    from multimodal_particles_b200.bridges import AbsorbingBridge
    from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
    from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
    from multimodal_particles_b200.states import MultiHeadOutput
    cfg = MultimodalBridgeMatchingConfig()
    torch.manual_seed(0)
    model = MultiModalBridgeMatching(cfg).to("cuda:0")
    g = torch.Generator().manual_seed(1)
    B, N, S = 4096, 128, 8
    mask = (torch.arange(N)[None] < torch.randint(1, N + 1, (B, 1), generator=g)).long().unsqueeze(-1)
    batch = SimpleNamespace(source_continuous=torch.randn(B, N, 3, generator=g), source_discrete=torch.randint(0, S, (B, N, 1), generator=g),
                            source_mask=mask, target_continuous=torch.randn(B, N, 3, generator=g) * mask,
                            target_discrete=torch.randint(0, S, (B, N, 1), generator=g) * mask, target_mask=mask)
    t = torch.rand(B, generator=g)
    t[:2] = torch.tensor([0.0, 1.0])
    state = model.sample_bridges(batch, t=t)
    assert state.time.shape == (B, 1, 1) and state.discrete.shape == (B, N, 1) and state.discrete.dtype == torch.int64
    kt = state.discrete.cpu()
    assert torch.equal(kt[0], batch.source_discrete[0]) and torch.equal(kt[1], batch.target_discrete[1])   # end points of the bridge
    sigma = cfg.bridge.sigma
    xt = state.continuous.cpu()
    lin = t.view(B, 1, 1) * batch.target_continuous + (1 - t.view(B, 1, 1)) * batch.source_continuous
    resid = (xt - lin) / sigma
    assert abs(float(resid.mean())) < 5e-3 and abs(float(resid.std()) - 1.0) < 5e-3
    stay = ((kt == batch.source_discrete) | (kt == batch.target_discrete)).float().mean()
    assert stay > 0.8                                   # posterior mass sits on the two end states for gamma = 0.125
    heads = model(state, batch)
    l0, l1 = model.loss_continuous(heads, state, batch), model.loss_discrete(heads, state, batch)
    v, lg = heads.continuous.cpu(), heads.discrete.cpu()
    m = mask.float()
    want0 = (((v - (batch.target_continuous - batch.source_continuous)) ** 2) * m).sum() / m.sum()
    want1 = (torch.nn.functional.cross_entropy(lg.reshape(-1, S), batch.target_discrete.reshape(-1), reduction="none") * m.reshape(-1)).sum() / m.sum()
    np.testing.assert_allclose([float(l0), float(l1)], [float(want0), float(want1)], rtol=2e-5)
    val = model.validation_step(batch, 0)
    assert val.ndim == 0 and torch.isfinite(val)
    with pytest.raises(NotImplementedError):
        model.training_step(batch, 0)
    from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
    acfg = AbsorbingConfig()
    acfg.data.max_num_particles = N
    flow = AbsorbingFlow(acfg).to("cuda:0")
    ast = flow.sample_bridges(batch, t=t)
    assert ast.mask_t.shape == (B, N, 1) and (ast.mask_t.cpu() >= mask).all() and ast.discrete.shape == (B, N, 1)
    assert torch.equal(ast.discrete[0].cpu(), batch.source_discrete[0]) and ast.continuous.shape == (B, N, 3)
    ab = AbsorbingBridge(AbsorbingConfig())
    mt = ab.sample(t.view(B, 1, 1).cuda(), mask.cuda())
    assert mt.shape == (B, N, 1) and (mt.cpu() >= mask).all() and torch.equal(mt[1].cpu(), mask[1])   # t = 1: only the targets survive


@pytest.mark.gpu
def test_absorbing_sample_philox_is_per_particle_for_any_width():
    """Default (in-kernel Philox) draws of AbsorbingBridge.sample: the scalar branch taken when N % 4 != 0 must use word
    n & 3 of the quad's Philox block exactly like the vectorised branch — particle n of jet b gets the same draw whatever
    the padded width is, and the four particles of a quad are independent (round-1 defect: they shared word .x)."""
    from multimodal_particles_b200 import _native
    B = 8192
    sp = torch.full((B,), 0.5, device="cuda:0")
    zeros = lambda n: torch.zeros(B, n, dtype=torch.uint8, device="cuda:0")
    m8 = _native.absorbing_sample(sp, zeros(8), None, seed=3, jet_offset=17).cpu().numpy()     # vector branch
    for n in (6, 7, 5, 3):                                                                          # scalar branch
        mn = _native.absorbing_sample(sp, zeros(n), None, seed=3, jet_offset=17).cpu().numpy()
        assert np.array_equal(mn, m8[:, :n]), n
    m6 = m8[:, :6].astype(np.float64)
    assert abs(m6.mean() - 0.5) < 0.01
    c = np.corrcoef(m6.T)
    assert np.abs(c - np.eye(6)).max() < 0.05, c     # sharing one uniform inside a quad gave correlation 1
