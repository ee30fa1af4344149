"""GPU: fused post-processing + jet observables (mmb_jet_observables) against the reference fixture and,
at the C2 batch size, against the oracle; plus the ParticleClouds / JetClassHighLevelFeatures mirror."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200 import HybridState
from multimodal_particles_b200.observables import JetClassHighLevelFeatures, ParticleClouds, jet_observables
from test_oracle_observables import GOLD, check_against_fixture

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_kernel_matches_reference_fixture():
    z = np.load(GOLD)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    x_phys, fc, jets = jet_observables(dev(z["x"]), dev(z["k"][..., 0]), dev(z["mask"][..., 0]), {"mean": z["mean"], "std": z["std"]})
    check_against_fixture(z, x_phys.cpu().numpy(), fc.cpu().numpy(), jets.cpu().numpy())


def test_mirror_classes_follow_the_reference_api():
    z = np.load(GOLD)
    state = HybridState(None, torch.from_numpy(z["x"]), torch.from_numpy(z["k"]).long(), torch.from_numpy(z["mask"]).long())
    pc = ParticleClouds(dataset=state)
    pc.postprocess(input_continuous="standardize", input_discrete="tokens", stats={"mean": z["mean"].tolist(), "std": z["std"].tolist()})
    assert pc.continuous.device.type == "cpu" and pc.discrete.shape == z["discrete"].shape
    np.testing.assert_allclose(pc.continuous.numpy(), z["continuous"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(pc.discrete.numpy(), z["discrete"]) and np.array_equal(pc.flavor.numpy(), z["flavor"])
    jets = JetClassHighLevelFeatures(pc)
    np.testing.assert_allclose(jets.pt.numpy(), z["jet_pt"], rtol=2e-5, atol=2e-5)
    assert np.array_equal(jets.multiplicity.numpy(), z["multiplicity"])
    pc.compute_4mom()
    np.testing.assert_allclose(pc.px.sum(-1).numpy(), z["jet_px"], rtol=1e-4, atol=1e-4)
    assert jets.Wassertein1D("pt", jets) == 0.0
    with pytest.raises(NotImplementedError):
        jets.substructure()


def test_full_batch_against_oracle():
    g = np.random.default_rng(11)
    B, N = 4096, 128
    mult = np.clip(np.rint(g.normal(45, 18, B)), 1, N).astype(int)
    mask = (np.arange(N)[None] < mult[:, None]).astype(np.uint8)
    x = (g.standard_normal((B, N, 3)) * mask[..., None]).astype(np.float32)
    k = (g.integers(0, 8, (B, N)) * mask).astype(np.uint8)
    stats = {"mean": [3.1, -0.02, 0.01], "std": [6.5, 0.21, 0.23]}
    wx, wfc, wj = ol.jet_observables(x, k, mask, stats)
    dev = lambda a: torch.from_numpy(a).to(DEV)
    gx, gfc, gj = jet_observables(dev(x), dev(k), dev(mask), stats)
    assert np.array_equal(gfc.cpu().numpy(), wfc)
    np.testing.assert_allclose(gx.cpu().numpy(), wx, rtol=1e-6, atol=1e-6)
    gj = gj.cpu().numpy()
    scale = np.abs(wj[:, 3]).max()
    for col in (0, 1, 2, 3, 4):
        np.testing.assert_allclose(gj[:, col], wj[:, col], rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(gj[:, 5] ** 2, wj[:, 5] ** 2, rtol=0, atol=3e-5 * scale ** 2)
    assert np.array_equal(gj[:, 8], wj[:, 8]) and np.array_equal(gj[:, 9], wj[:, 9])
    np.testing.assert_allclose(gj[:, 7], wj[:, 7], rtol=1e-4, atol=1e-4)
