"""Context features (reference: multimodal_bridge_matching.py:143-144, architectures/utils.py:84-96,155-170, architectures/epic.py:
187-189,226-238): the per-jet context vector [time embedding | embedded context] enters global_0, fc_global1 and fc_local1.
CPU: the oracle and the host mirror against fixtures produced by the unmodified reference with context features switched on
(tests/golden/make_golden_context.py).  GPU: fp32 kernel bit-exact against the oracle, warp-MMA engine against fp32, host paths."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from test_oracle_golden import HEAD_ATOL, HEAD_RTOL, _near_threshold

CTX_CASES = ["mbm_ctx", "mbm_ctx_id"]
DEV = "cuda:0"


def load(golden_dir, case):
    z, cfg, model = ol.load_mbm_golden(os.path.join(golden_dir, case + ".npz"))
    dims, packed = ol.packed_model(model)
    ctx = model.encoder.epic.embedding.context(torch.from_numpy(z["context_continuous"]), None, "cpu").numpy()
    return z, cfg, model, dims, packed, ctx


def rows(temb_i, ctx):
    """[B, T + X]: the step's time embedding followed by each jet's embedded context."""
    return np.concatenate([np.repeat(temb_i[None], len(ctx), 0), ctx], axis=1).astype(np.float32)


@pytest.mark.parametrize("case", CTX_CASES)
def test_dims_and_state_dict(case, golden_dir):
    z, cfg, model, dims, packed, ctx = load(golden_dir, case)
    want = cfg.encoder.dim_emb_context_continuous or cfg.data.dim_context_continuous
    assert dims.dim_context == want == ctx.shape[1] and dims.dim_time_emb == cfg.encoder.dim_emb_time
    assert ol.lib().mmbo_expf  # oracle loads
    H, G, T, X = dims.dim_hidden_local, dims.dim_hidden_glob, dims.dim_time_emb, dims.dim_context
    assert model.encoder.epic.epic.epic_proj.global_0.weight_v.shape == (H, 2 * H + T + X)
    assert model.encoder.epic.epic.epic_layers[0].fc_local1.weight_v.shape == (H, H + G + T + X)


@pytest.mark.parametrize("case", CTX_CASES)
def test_network_heads_match_reference(case, golden_dir):
    z, cfg, model, dims, packed, ctx = load(golden_dir, case)
    for i in z["snap_steps"]:
        v, logits = ol.epic_forward(dims, packed, z[f"snap{i}/x"], z[f"snap{i}/k"], z["mask"], rows(z["temb"][i], ctx))
        np.testing.assert_allclose(v, z[f"snap{i}/v"], rtol=HEAD_RTOL, atol=HEAD_ATOL)
        np.testing.assert_allclose(logits, z[f"snap{i}/logits"], rtol=HEAD_RTOL, atol=HEAD_ATOL)
    # the context matters: another jet's context gives other heads
    i = int(z["snap_steps"][0])
    v2, _ = ol.epic_forward(dims, packed, z[f"snap{i}/x"], z[f"snap{i}/k"], z["mask"], rows(z["temb"][i], np.roll(ctx, 1, 0)))
    assert np.abs(v2 - z[f"snap{i}/v"]).max() > 1e-3


@pytest.mark.parametrize("case", CTX_CASES)
def test_generation_matches_reference_trajectory(case, golden_dir):
    """Whole simulate_dynamics with context: jets that follow the reference's tokens end within fp32 drift of its features;
    a jet may leave the trajectory only through a draw on a categorical threshold (as in test_oracle_golden)."""
    z, cfg, model, dims, packed, ctx = load(golden_dir, case)
    tab = model.step_table()
    x, k = ol.generate(dims, packed, z["x0"], z["k0"][..., 0], z["mask"][..., 0], tab, u_jump=z["u_jump"], context=ctx)
    same = (k == z["k_final"][..., 0]).all(-1)
    assert same.mean() >= 0.6, f"only {same.sum()}/{len(same)} jets reproduce the reference tokens"
    np.testing.assert_allclose(x[same], z["x_final"][same], rtol=1e-4, atol=1e-4)
    xs, ks, mask = z["x0"].copy(), z["k0"][..., 0].copy(), z["mask"][..., 0]
    alive, found = np.ones(len(xs), bool), 0
    for s in range(tab.n_steps):
        v, logits = ol.epic_forward(dims, packed, xs, ks, mask, rows(tab.temb[s].numpy(), ctx))
        before = ks
        xs, ks, _ = ol.bridge_update(xs, ks, mask, v, logits, z["u_jump"][s], tab.dt, float(tab.bc[s]), float(tab.cc[s]))
        diff = (ks != z["k_traj"][s]) & alive[:, None]
        if diff.any():
            near = _near_threshold(logits, before, z["u_jump"][s], tab.dt, float(tab.bc[s]), float(tab.cc[s]), tol=2e-4)
            assert not (diff & ~near).any(), f"step {s}: a jet left the reference trajectory through a draw that is not on a threshold"
            found += int(diff.any(-1).sum())
            alive &= ~diff.any(-1)
    assert found >= int((~same).sum())
    assert np.array_equal(xs, x) and np.array_equal(ks, k)   # stepwise == fused oracle loop, with context


def test_missing_context_is_an_error_and_discrete_context_is_refused(golden_dir):
    from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
    from multimodal_particles_b200.multimodal_bridge_matching import MultiModalBridgeMatching
    z, cfg, model, dims, packed, ctx = load(golden_dir, "mbm_ctx")
    with pytest.raises(ValueError):
        model.encoder.epic.embedding.context(None, None, "cpu")
    cfg = MultimodalBridgeMatchingConfig()
    cfg.data.dim_context_discrete, cfg.data.vocab_size_context = 1, 3
    cfg.encoder.embedding_context_discrete, cfg.encoder.dim_emb_context_discrete = "Embedding", 4
    with pytest.raises(NotImplementedError):   # the reference's own forward fails on such a model (utils.py:100 vs :161)
        MultiModalBridgeMatching(cfg)
    # the absorbing generator cannot take context in the reference either: absorbing_flows.py:150 wraps it in a tuple, the
    # wrapper drops the non-tensor and the context embedding is applied to None
    from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
    from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
    acfg = AbsorbingConfig()
    acfg.data.dim_context_continuous = 2
    acfg.encoder.embedding_context_continuous, acfg.encoder.dim_emb_context_continuous = "Embedding", 3
    with pytest.raises(NotImplementedError):
        AbsorbingFlow(acfg)


# ---- GPU --------------------------------------------------------------------------------------------------------------------
def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(DEV, dtype) if dtype else t.to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CTX_CASES)
def test_gpu_forward_fp32_bit_exact(case, golden_dir):
    z, cfg, model, dims, packed, ctx = load(golden_dir, case)
    native = model.to(DEV).encoder.native_model(torch.device(DEV))
    for i in z["snap_steps"]:
        x, k, mask = z[f"snap{i}/x"], z[f"snap{i}/k"][..., 0], z["mask"][..., 0]
        r = rows(z["temb"][i], ctx)
        vw, lw = ol.epic_forward(dims, packed, x, k, mask, r)
        v, lg = native.forward(dev(x), dev(k), dev(mask), dev(r), precision="fp32")
        assert np.array_equal(v.cpu().numpy().view(np.uint32), vw.view(np.uint32))
        assert np.array_equal(lg.cpu().numpy().view(np.uint32), lw.view(np.uint32))
        np.testing.assert_allclose(lg.cpu().numpy(), z[f"snap{i}/logits"], rtol=HEAD_RTOL, atol=HEAD_ATOL)
    with pytest.raises(Exception):   # the tcgen05 trunk has no context path
        native.forward(dev(x), dev(k), dev(mask), dev(r), precision="bf16")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CTX_CASES)
def test_gpu_generation_fp32_bit_exact_and_mirror(case, golden_dir):
    """mmb_generate (fp32, injected uniforms and in-kernel Philox) == oracle bit for bit; MultiModalBridgeMatching.forward and
    simulate_dynamics read the context from the batch like the reference."""
    from types import SimpleNamespace
    from multimodal_particles_b200 import HybridState
    z, cfg, model, dims, packed, ctx = load(golden_dir, case)
    model = model.to(DEV)
    native = model.encoder.native_model(torch.device(DEV))
    tab = model.step_table()
    for u in (z["u_jump"], None):
        xw, kw = ol.generate(dims, packed, z["x0"], z["k0"][..., 0], z["mask"][..., 0], tab, u_jump=u, seed=5, jet_offset=11, context=ctx)
        x, k = dev(z["x0"]).clone(), dev(z["k0"][..., 0]).clone()
        native.generate(x, k, dev(z["mask"][..., 0]), tab, u_jump=None if u is None else dev(u), seed=5, jet_offset=11,
                        precision="fp32", context=dev(ctx))
        assert np.array_equal(x.cpu().numpy().view(np.uint32), xw.view(np.uint32)) and np.array_equal(k.cpu().numpy(), kw)
    assert native.generate_precision(z["x0"].shape[1]) in ("f16", "fp32")   # never the tcgen05 engine
    with pytest.raises(Exception):
        native.generate(x, k, dev(z["mask"][..., 0]), tab, precision="fp32")   # context missing
    batch = SimpleNamespace(context_continuous=torch.from_numpy(z["context_continuous"]))
    state = lambda: HybridState(None, torch.from_numpy(z["x0"]).clone(), torch.from_numpy(z["k0"]).long(), torch.from_numpy(z["mask"]).long())
    out = model.simulate_dynamics(state(), batch, uniforms=torch.from_numpy(z["u_jump"]), precision="fp32")
    xw, kw = ol.generate(dims, packed, z["x0"], z["k0"][..., 0], z["mask"][..., 0], tab, u_jump=z["u_jump"], context=ctx)
    assert np.array_equal(out.continuous.numpy(), xw) and np.array_equal(out.discrete[..., 0].numpy(), kw)
    i = int(z["snap_steps"][1])
    st = HybridState(torch.full((len(ctx), 1, 1), float(z["t"][i])), dev(z[f"snap{i}/x"]), dev(z[f"snap{i}/k"]).long(), dev(z["mask"]).long())
    heads = model(st, batch)
    np.testing.assert_allclose(heads.continuous.cpu().numpy(), z[f"snap{i}/v"], rtol=HEAD_RTOL, atol=HEAD_ATOL)
    np.testing.assert_allclose(heads.discrete.cpu().numpy(), z[f"snap{i}/logits"], rtol=HEAD_RTOL, atol=HEAD_ATOL)


@pytest.mark.gpu
def test_gpu_generation_f16_with_context_tracks_fp32_and_host_paths_agree(golden_dir):
    """The warp-MMA engine with per-jet context terms: 2048 jets with the fixture's weights and random contexts against the
    fp32 kernel (same Philox draws): token agreement >= 0.97, mean |dx| <= 0.01 (the bars of the context-free engine,
    tests/test_gpu_mma.py), and far closer to fp32 than a run with the contexts permuted; the host entry point
    (direct mode and sliced mode) returns the same bits as the device call; a wrong context changes the result."""
    from types import SimpleNamespace
    from multimodal_particles_b200 import HybridState
    from multimodal_particles_b200.databatch import jetclass_like_databatch
    z, cfg, model, dims, packed, ctx = load(golden_dir, "mbm_ctx")
    model = model.to(DEV)
    native = model.encoder.native_model(torch.device(DEV))
    assert native.generate_precision(128) == "f16"
    B = 2048
    b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(77))
    cc = torch.randn(B, z["context_continuous"].shape[1], generator=torch.Generator().manual_seed(78)) * 1.5
    batch = SimpleNamespace(context_continuous=cc)
    mk = lambda pin=False: HybridState(None, *[(t.clone().pin_memory() if pin else t.clone().to(DEV))
                                               for t in (b.source_continuous, b.source_discrete, b.source_mask)])
    model.seed = 3
    ref = model.simulate_dynamics(mk(), batch, precision="fp32", jet_offset=100)
    got = model.simulate_dynamics(mk(), batch, precision="f16", jet_offset=100)
    live = b.source_mask[..., 0].bool()
    agree = (ref.discrete[..., 0] == got.discrete[..., 0])[live].float().mean().item()
    assert agree >= 0.97, agree
    dx = (ref.continuous - got.continuous).abs()[live].mean().item()
    assert dx <= 0.01, dx
    for chunks in (0, 2):
        model.pipeline_chunks, model.pipeline_min_jets = chunks, 1
        host = model.simulate_dynamics(mk(pin=True), batch, precision="f16", jet_offset=100)
        assert torch.equal(host.continuous, got.continuous) and torch.equal(host.discrete, got.discrete)
    other = model.simulate_dynamics(mk(), SimpleNamespace(context_continuous=cc.roll(1, 0)), precision="f16", jet_offset=100)
    assert (other.continuous - got.continuous).abs()[live].mean().item() > 10 * dx
