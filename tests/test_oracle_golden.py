"""CPU: the C oracle and the host logic against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200.steptable import build_step_table

MBM_CASES = ["mbm_c1", "mbm_n128", "mbm_odd", "mbm_wide"]   # mbm_wide: the class-default widths (epic.py:99-101)
# fp32 network tolerance vs torch's CPU GEMM ordering
HEAD_RTOL, HEAD_ATOL = 2e-5, 2e-5


def test_expf_accuracy():
    xs = np.concatenate([np.linspace(-87, 88, 200001), np.linspace(-1, 1, 20001)]).astype(np.float32)
    got = np.array([ol.lib().mmbo_expf(float(x)) for x in xs[::7]], dtype=np.float64)
    want = np.exp(xs[::7].astype(np.float64))
    rel = np.abs(got - want) / want
    assert rel.max() < 2.5e-7  # ~2 ulp
    assert ol.lib().mmbo_expf(-90.0) == 0.0 and ol.lib().mmbo_expf(0.0) == 1.0


@pytest.mark.parametrize("case", MBM_CASES)
def test_step_table_equals_reference(case, golden_dir):
    """Host step table == the scalars the reference itself used, bit for bit."""
    z, cfg, _ = ol.load_mbm_golden(os.path.join(golden_dir, case + ".npz"))
    tab = build_step_table(cfg.bridge.num_timesteps, cfg.bridge.time_eps, cfg.data.vocab_size_features,
                           cfg.bridge.gamma, cfg.encoder.dim_emb_time)
    assert tab.n_steps == len(z["t"])
    assert np.array_equal(tab.t.numpy(), z["t"])
    assert np.array_equal(tab.temb.numpy(), z["temb"])
    assert np.array_equal(tab.bc.numpy(), z["bc"]) and np.array_equal(tab.cc.numpy(), z["cc"])
    assert np.float32(tab.dt) == z["dt"]
    # the oracle's standalone libm table agrees to rounding (it is not used for parity runs)
    lm = ol.step_table_libm(cfg.bridge.num_timesteps, cfg.bridge.time_eps, cfg.data.vocab_size_features,
                            cfg.bridge.gamma, cfg.encoder.dim_emb_time)
    np.testing.assert_allclose(lm.t, z["t"], rtol=3e-7)
    np.testing.assert_allclose(lm.temb, z["temb"], atol=3e-7)
    np.testing.assert_allclose(lm.bc[:-3], z["bc"][:-3], rtol=2e-5)
    np.testing.assert_allclose(lm.bc[-3:], z["bc"][-3:], rtol=5e-3)  # (1-w) cancels at t -> 1


@pytest.mark.parametrize("case", MBM_CASES)
def test_network_heads_match_reference(case, golden_dir):
    """Oracle EPiC forward on the reference's own intermediate states == the reference's heads."""
    z, cfg, model = ol.load_mbm_golden(os.path.join(golden_dir, case + ".npz"))
    dims, packed = ol.packed_model(model)
    for i in z["snap_steps"]:
        x, k = z[f"snap{i}/x"], z[f"snap{i}/k"]
        v, logits = ol.epic_forward(dims, packed, x, k, z["mask"], z["temb"][i][None])
        np.testing.assert_allclose(v, z[f"snap{i}/v"], rtol=HEAD_RTOL, atol=HEAD_ATOL)
        np.testing.assert_allclose(logits, z[f"snap{i}/logits"], rtol=HEAD_RTOL, atol=HEAD_ATOL)


def _near_threshold(logits, k, u, dt, bc, cc, tol=1e-6):
    """Draws whose uniform lies within `tol` (relative) of a categorical threshold, in fp64."""
    l = logits.astype(np.float64)
    q = np.exp(l - l.max(-1, keepdims=True))
    q /= q.sum(-1, keepdims=True)
    qk = np.take_along_axis(q, k[..., None].astype(np.int64), -1)
    lam = (1.0 + bc * q + cc * qk) * dt
    c = np.cumsum(lam * np.exp(-lam.sum(-1, keepdims=True)), -1)
    return (np.abs(u[..., None] - c) <= tol * np.maximum(c, 1e-30)).any(-1)


def test_bridge_update_matches_reference_solver_steps(golden_dir):
    """Fused update == AbsorbingBridge/LinearUniformBridge/TelegraphBridge.solver_step of the
    reference on identical heads and injected uniforms: tokens and masks exact, x exact."""
    z = np.load(os.path.join(golden_dir, "bridge_update.npz"))
    dt = float(z["dt"])
    n_checked = n_near = 0
    for i in z["steps"]:
        g = lambda name: z[f"s{i}/{name}"]
        x, k, mask = g("in/x"), g("in/k")[..., 0], g("in/mask")[..., 0]
        bc, cc, sp = float(g("bc")), float(g("cc")), float(g("sp"))
        near = _near_threshold(g("in/logits"), k, g("in/uj"), dt, bc, cc)
        # multimodal
        x1, k1, _ = ol.bridge_update(x, k, mask, g("in/v"), g("in/logits"), g("in/uj"), dt, bc, cc)
        assert np.array_equal(x1, g("mbm/x"))
        bad = (k1 != g("mbm/k")[..., 0])
        assert not (bad & ~near).any(), f"step {i}: {int((bad & ~near).sum())} unexplained token mismatches"
        # absorbing flow ordering: birth, Euler with new mask, jump with new mask
        x2, k2, m2 = ol.bridge_update(x, k, mask, g("in/v"), g("in/logits"), g("in/uj"), dt, bc, cc,
                                      absorb_logit=g("in/a")[..., 0], u_absorb=g("in/ua"), sp=sp, flags=1)
        assert np.array_equal(m2, g("abs/mask")[..., 0])
        assert np.array_equal(x2, g("abs/x"))
        bad2 = (k2 != g("abs/k")[..., 0])
        assert not (bad2 & ~near).any()
        n_checked += k.size
        n_near += int(near.sum())
    assert n_near <= 2, f"{n_near} near-threshold draws in {n_checked}: more than rounding can explain"


@pytest.mark.parametrize("case", MBM_CASES)
def test_generation_matches_reference_trajectory(case, golden_dir):
    """Whole simulate_dynamics: oracle trajectory vs the reference's with the same uniforms.
    A jet whose tokens agree at every step must end within fp32 drift of the reference.  A jet may diverge only through a
    draw that sits on a categorical threshold: for every diverging jet the FIRST differing (step, particle) is located in the
    reference's recorded token trajectory and the oracle's logits at that step must put the uniform within 2e-4 (relative)
    of a threshold — the size of the head differences between the two implementations (2e-5 on logits of magnitude ~10)."""
    z, cfg, model = ol.load_mbm_golden(os.path.join(golden_dir, case + ".npz"))
    dims, packed = ol.packed_model(model)
    tab = model.step_table()
    x, k = ol.generate(dims, packed, z["x0"], z["k0"][..., 0], z["mask"][..., 0], tab, u_jump=z["u_jump"])
    same = (k == z["k_final"][..., 0]).all(-1)
    assert same.mean() >= 0.75, f"only {same.sum()}/{len(same)} jets reproduce the reference tokens"
    np.testing.assert_allclose(x[same], z["x_final"][same], rtol=1e-4, atol=1e-4)
    explained = first_divergences_are_threshold_draws(z, dims, packed, tab)
    assert explained >= int((~same).sum())   # every jet that ends elsewhere left the trajectory at an explained draw


def first_divergences_are_threshold_draws(z, dims, packed, tab, tol=2e-4):
    """Step the oracle (epic_forward + bridge_update == generate, see the next test) beside the reference's token trajectory;
    returns the number of jets whose first divergence was found, asserting each one is a near-threshold draw."""
    xs, ks, mask = z["x0"].copy(), z["k0"][..., 0].copy(), z["mask"][..., 0]
    ref_traj = z["k_traj"]
    alive = np.ones(len(xs), bool)      # jets that still follow the reference
    found = 0
    for s in range(tab.n_steps):
        v, logits = ol.epic_forward(dims, packed, xs, ks, mask, tab.temb[s].numpy()[None])
        k_before = ks
        xs, ks, _ = ol.bridge_update(xs, ks, mask, v, logits, z["u_jump"][s], tab.dt, float(tab.bc[s]), float(tab.cc[s]))
        diff = (ks != ref_traj[s]) & alive[:, None]
        if diff.any():
            near = _near_threshold(logits, k_before, z["u_jump"][s], tab.dt, float(tab.bc[s]), float(tab.cc[s]), tol=tol)
            assert not (diff & ~near).any(), f"step {s}: a jet left the reference trajectory through a draw that is not on a threshold"
            newly = diff.any(-1)
            found += int(newly.sum())
            alive &= ~newly
    return found


def test_generation_stepwise_equals_fused_oracle(golden_dir):
    """oracle.generate == repeated (epic_forward, bridge_update): the fused loop adds nothing."""
    z, cfg, model = ol.load_mbm_golden(os.path.join(golden_dir, "mbm_c1.npz"))
    dims, packed = ol.packed_model(model)
    tab = model.step_table()
    x, k, mask = z["x0"].copy(), z["k0"][..., 0].copy(), z["mask"][..., 0]
    for s in range(tab.n_steps):
        v, logits = ol.epic_forward(dims, packed, x, k, mask, tab.temb[s].numpy()[None])
        x, k, _ = ol.bridge_update(x, k, mask, v, logits, z["u_jump"][s], tab.dt, float(tab.bc[s]), float(tab.cc[s]))
    xg, kg = ol.generate(dims, packed, z["x0"], z["k0"][..., 0], mask, tab, u_jump=z["u_jump"])
    assert np.array_equal(x, xg) and np.array_equal(k, kg)
    # token trajectory of the reference, step by step, for the jets that never diverge
    assert kg.shape == z["k_traj"][-1].shape


def test_philox_uniforms_properties():
    u = ol.philox_uniforms(seed=7, jet_offset=0, n_steps=3, B=5, N=30)
    assert u.min() >= 0.0 and u.max() < 1.0
    # sharding invariance: jets [2,5) drawn with jet_offset=2 equal the slice of the full draw
    u2 = ol.philox_uniforms(seed=7, jet_offset=2, n_steps=3, B=3, N=30)
    assert np.array_equal(u[:, 2:], u2)
    big = ol.philox_uniforms(seed=1, jet_offset=0, n_steps=4, B=64, N=128)
    assert abs(big.mean() - 0.5) < 5e-3 and abs(big.var() - 1 / 12) < 2e-3
    assert len(np.unique(big)) > 0.99 * big.size * (1 - big.size / 2**25)
