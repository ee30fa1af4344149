"""GPU parity: libmmbridge.so (through the C ABI / the Python mirror) against the CPU oracle on the
same seeded inputs, against the committed golden fixtures of the reference, and — at the full
BASELINE sizes — through size-independent properties.

Bars: tokens, masks and every fp32-path output are BIT-EXACT against the oracle; the bf16
(tcgen05) trunk is checked against the fp32 path within the tolerance written at each test.
"""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200 import HybridState, MultiModalBridgeMatching, _native
from multimodal_particles_b200.bridges import NO_EULER, NO_JUMP
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.databatch import jetclass_like_databatch, random_databatch
from multimodal_particles_b200.states import MultiHeadOutput

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(DEV, dtype) if dtype else t.to(DEV)


def rand_update_inputs(B, N, Dc, S, seed, logit_scale=3.0):
    g = np.random.default_rng(seed)
    return dict(
        x=g.standard_normal((B, N, Dc), dtype=np.float32) * 2,
        k=g.integers(0, S, (B, N)).astype(np.uint8),
        mask=g.integers(0, 2, (B, N)).astype(np.uint8),
        v=g.standard_normal((B, N, Dc), dtype=np.float32) * 3,
        logits=g.standard_normal((B, N, S), dtype=np.float32) * logit_scale,
        a=g.standard_normal((B, N), dtype=np.float32) * 3,
        uj=(g.random((B, N), dtype=np.float32) * np.where(g.random((B, N)) < 0.3, 0.02, 1.0)).astype(np.float32),
        ua=g.random((B, N), dtype=np.float32),
    )


def cuda_update(d, dt, bc, cc, sp=0.0, flags=0):
    x, k, m = dev(d["x"]).clone(), dev(d["k"]).clone(), dev(d["mask"]).clone()
    absorbing = bool(flags & 1)
    _native.bridge_update(x, k, m, dev(d["v"]), dev(d["logits"]), dev(d["uj"]), dt, bc, cc,
                          absorb_logit=dev(d["a"]) if absorbing else None,
                          u_absorb=dev(d["ua"]) if absorbing else None, sp=sp, flags=flags)
    torch.cuda.synchronize()
    return x.cpu().numpy(), k.cpu().numpy(), m.cpu().numpy()


# late steps have bc up to ~8e4; cover early, late and last-step coefficients
COEFFS = [(0.0101, 4.7, 0.37, 0.93), (0.0101, 51.3, 0.865, 0.12), (0.0101, 79991.0, 0.9999, 1.2e-5)]


@pytest.mark.parametrize("B,N,Dc,S", [(64, 30, 3, 4), (32, 128, 3, 8), (7, 37, 2, 5), (3, 33, 3, 8), (1, 1, 3, 8)])
@pytest.mark.parametrize("flags", [0, 1, NO_EULER, NO_JUMP, 1 | NO_EULER | NO_JUMP])
def test_bridge_update_bit_exact_vs_oracle(B, N, Dc, S, flags):
    d = rand_update_inputs(B, N, Dc, S, seed=B * 131 + N + flags)
    for dt, bc, cc, sp in COEFFS:
        want = ol.bridge_update(d["x"], d["k"], d["mask"], d["v"], d["logits"], d["uj"], dt, bc, cc,
                                absorb_logit=d["a"], u_absorb=d["ua"], sp=sp, flags=flags)
        got = cuda_update(d, dt, bc, cc, sp, flags)
        for w, g_, name in zip(want, got, "xkm"):
            assert np.array_equal(w.view(np.uint8) if w.dtype == np.float32 else w,
                                  g_.view(np.uint8) if g_.dtype == np.float32 else g_), f"{name} differs (flags={flags})"


def test_bridge_update_full_size_c2():
    """BASELINE config 2 size (4096 x 128, S=8): exact against the oracle, plus invariants."""
    d = rand_update_inputs(4096, 128, 3, 8, seed=2)
    dt, bc, cc, _ = COEFFS[1]
    xw, kw, _ = ol.bridge_update(d["x"], d["k"], d["mask"], d["v"], d["logits"], d["uj"], dt, bc, cc)
    xg, kg, mg = cuda_update(d, dt, bc, cc)
    assert np.array_equal(xw.view(np.uint32), xg.view(np.uint32)) and np.array_equal(kw, kg)
    dead = d["mask"] == 0
    assert (kg[dead] == 0).all() and (xg[dead] == 0).all() and kg.max() < 8
    assert np.array_equal(mg, d["mask"])


def test_bridge_update_vs_reference_fixture(golden_dir):
    """C-ABI kernel against the reference's own solver_steps (tests/golden/bridge_update.npz)."""
    z = np.load(os.path.join(golden_dir, "bridge_update.npz"))
    dt = float(z["dt"])
    mismatches = 0
    for i in z["steps"]:
        g = lambda name: z[f"s{i}/{name}"]
        d = dict(x=g("in/x"), k=g("in/k")[..., 0], mask=g("in/mask")[..., 0], v=g("in/v"), logits=g("in/logits"),
                 a=g("in/a")[..., 0], uj=g("in/uj"), ua=g("in/ua"))
        bc, cc, sp = float(g("bc")), float(g("cc")), float(g("sp"))
        x1, k1, _ = cuda_update(d, dt, bc, cc)
        assert np.array_equal(x1, g("mbm/x"))
        mismatches += int((k1 != g("mbm/k")[..., 0]).sum())
        x2, k2, m2 = cuda_update(d, dt, bc, cc, sp, flags=1)
        assert np.array_equal(m2, g("abs/mask")[..., 0]) and np.array_equal(x2, g("abs/x"))
        mismatches += int((k2 != g("abs/k")[..., 0]).sum())
    assert mismatches <= 2, "token mismatches beyond near-threshold rounding (see test_oracle_golden)"


def test_philox_uniforms_bit_exact():
    u = _native.philox_uniforms(1234, 77, 5, 9, 30, DEV).cpu().numpy()
    assert np.array_equal(u, ol.philox_uniforms(1234, 77, 5, 9, 30))


# ---------------------------------------------------------------------------------------------
def golden_model(golden_dir, case):
    z, cfg, model = ol.load_mbm_golden(os.path.join(golden_dir, case + ".npz"))
    dims, packed = ol.packed_model(model)
    return z, cfg, model.to(DEV), dims, packed


@pytest.mark.parametrize("case", ["mbm_c1", "mbm_n128", "mbm_odd"])
def test_epic_forward_fp32_bit_exact_and_matches_reference(case, golden_dir):
    z, cfg, model, dims, packed = golden_model(golden_dir, case)
    native = model.encoder.native_model(torch.device(DEV))
    for i in z["snap_steps"]:
        x, k, mask = z[f"snap{i}/x"], z[f"snap{i}/k"][..., 0], z["mask"][..., 0]
        temb = z["temb"][i][None]
        v, lg, hid = native.forward(dev(x), dev(k), dev(mask), dev(temb), want_hidden=True, precision="fp32")
        vo, lo, ho = ol.epic_forward(dims, packed, x, k, mask, temb, want_hidden=True)
        assert np.array_equal(v.cpu().numpy(), vo) and np.array_equal(lg.cpu().numpy(), lo)
        assert np.array_equal(hid.cpu().numpy(), ho)
        # and the reference's heads, within fp32 GEMM-ordering tolerance
        np.testing.assert_allclose(v.cpu().numpy(), z[f"snap{i}/v"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(lg.cpu().numpy(), z[f"snap{i}/logits"], rtol=2e-5, atol=2e-5)


def test_epic_forward_per_jet_times(golden_dir):
    """temb_stride = T: every jet at its own time (training-style call of model(state, batch))."""
    z, cfg, model, dims, packed = golden_model(golden_dir, "mbm_c1")
    B = z["x0"].shape[0]
    t = torch.linspace(0.05, 0.95, B).reshape(B, 1, 1)
    state = HybridState(t.to(DEV), dev(z["x0"]), dev(z["k0"]).long(), dev(z["mask"]).long())
    heads = model(state, None)
    temb = model.encoder.epic.time_embedding(t).numpy()
    vo, lo = ol.epic_forward(dims, packed, z["x0"], z["k0"][..., 0], z["mask"][..., 0], temb)
    assert np.array_equal(heads.continuous.cpu().numpy(), vo) and np.array_equal(heads.discrete.cpu().numpy(), lo)
    assert heads.absorbing is state.absorbing


@pytest.mark.parametrize("case", ["mbm_c1", "mbm_n128", "mbm_odd"])
def test_generate_fp32_bit_exact_and_matches_reference(case, golden_dir):
    """simulate_dynamics through the public API == oracle bit for bit == reference fixture."""
    z, cfg, model, dims, packed = golden_model(golden_dir, case)
    state = HybridState(None, torch.from_numpy(z["x0"]), torch.from_numpy(z["k0"]).long(), torch.from_numpy(z["mask"]).long())
    out = model.simulate_dynamics(state, (z["x0"],), uniforms=torch.from_numpy(z["u_jump"]), precision="fp32")
    assert out.continuous.device.type == "cpu" and out.discrete.dtype == torch.int64 and out.discrete.shape[-1] == 1
    xo, ko = ol.generate(dims, packed, z["x0"], z["k0"][..., 0], z["mask"][..., 0], model.step_table(), u_jump=z["u_jump"])
    assert np.array_equal(out.continuous.numpy(), xo) and np.array_equal(out.discrete[..., 0].numpy(), ko)
    same = (ko == z["k_final"][..., 0]).all(-1)
    assert same.mean() >= 0.75
    np.testing.assert_allclose(xo[same], z["x_final"][same], rtol=1e-4, atol=1e-4)


def test_generate_philox_mode_and_sharding_invariance(golden_dir):
    """In-kernel Philox draws == oracle's; generating a batch in two shards (jet_offset) == in one go."""
    z, cfg, model, dims, packed = golden_model(golden_dir, "mbm_c1")
    x0, k0, mask = z["x0"], z["k0"][..., 0], z["mask"][..., 0]
    native = model.encoder.native_model(torch.device(DEV))
    tab = model.step_table()
    x, k = dev(x0).clone(), dev(k0).clone()
    native.generate(x, k, dev(mask), tab, seed=99, jet_offset=10, precision="fp32")
    xo, ko = ol.generate(dims, packed, x0, k0, mask, tab, seed=99, jet_offset=10)
    assert np.array_equal(x.cpu().numpy(), xo) and np.array_equal(k.cpu().numpy(), ko)
    xa, ka = dev(x0[:2]).clone(), dev(k0[:2]).clone()
    xb, kb = dev(x0[2:]).clone(), dev(k0[2:]).clone()
    native.generate(xa, ka, dev(mask[:2]), tab, seed=99, jet_offset=10, precision="fp32")
    native.generate(xb, kb, dev(mask[2:]), tab, seed=99, jet_offset=12, precision="fp32")
    assert torch.equal(torch.cat([xa, xb]), x) and torch.equal(torch.cat([ka, kb]), k)


def test_solver_steps_match_reference_fixture(golden_dir):
    """bridge.solver_step(state, heads, dt) one by one, as reference callers use them."""
    z = np.load(os.path.join(golden_dir, "bridge_update.npz"))
    cfg = MultimodalBridgeMatchingConfig()
    cfg.bridge.num_timesteps = 100
    model = MultiModalBridgeMatching(cfg)
    i = int(z["steps"][1])
    g = lambda name: z[f"s{i}/{name}"]
    B = g("in/x").shape[0]
    state = HybridState(torch.full((B, 1), float(g("t")), device=DEV), dev(g("in/x")), dev(g("in/k")).long(),
                        dev(g("in/mask")).long())
    heads = MultiHeadOutput(dev(g("in/v")), dev(g("in/logits")), state.absorbing)
    dt = torch.tensor(float(z["dt"]))
    x_alias = state.continuous
    state = model.bridge_continuous.solver_step(state, heads, dt)
    state = model.bridge_discrete.solver_step(state, heads, dt, uniforms=dev(g("in/uj")))
    assert state.continuous is x_alias  # mutated in place like the reference
    assert np.array_equal(state.continuous.cpu().numpy(), g("mbm/x"))
    assert (state.discrete.cpu().numpy()[..., 0] != g("mbm/k")[..., 0]).sum() <= 1
    assert state.discrete.dtype == torch.int64


def test_out_of_range_tokens_assert():
    cfg = MultimodalBridgeMatchingConfig()
    cfg.bridge.num_timesteps = 5
    model = MultiModalBridgeMatching(cfg).to(DEV)
    b = random_databatch(cfg)
    bad = b.source_discrete.clone()
    bad[0, 0, 0] = cfg.data.vocab_size_features
    with pytest.raises(AssertionError):
        model.simulate_dynamics(HybridState(None, b.source_continuous, bad, b.source_mask), b, precision="fp32")


def test_empty_jet_gives_nan_like_reference():
    """mean pooling divides by mask.sum (epic.py:141): a jet without particles is NaN, others fine."""
    cfg = MultimodalBridgeMatchingConfig()
    model = MultiModalBridgeMatching(cfg).to(DEV)
    g = torch.Generator().manual_seed(3)
    b = jetclass_like_databatch(4, generator=g)
    mask = b.source_mask.clone()
    mask[1] = 0
    state = HybridState(torch.full((4, 1), 0.3), b.source_continuous.to(DEV), b.source_discrete.to(DEV), mask.to(DEV))
    state.time = state.time.to(DEV)
    heads = model(state, b)
    assert torch.isfinite(heads.continuous[[0, 2, 3]]).all() and torch.isfinite(heads.discrete[[0, 2, 3]]).all()
    assert torch.isnan(heads.discrete[1]).all()


def test_full_size_c2_properties():
    """BASELINE config 2 (B=4096, N=128, S=8, 99 steps): determinism, shard invariance, mask
    invariants, and exactness against the oracle on a slice the CPU finishes in seconds."""
    cfg = MultimodalBridgeMatchingConfig()
    cfg.bridge.num_timesteps = 100
    torch.manual_seed(0)
    model = MultiModalBridgeMatching(cfg).to(DEV)
    g = torch.Generator().manual_seed(1234)
    b = jetclass_like_databatch(4096, generator=g)
    mk = lambda sl=slice(None): HybridState(None, b.source_continuous[sl].clone(), b.source_discrete[sl].clone(),
                                            b.source_mask[sl].clone())
    model.seed = 42
    out = model.simulate_dynamics(mk(), b, precision="fp32", jet_offset=0)
    again = model.simulate_dynamics(mk(), b, precision="fp32", jet_offset=0)
    assert torch.equal(out.continuous, again.continuous) and torch.equal(out.discrete, again.discrete)
    # the host -> host call above ran as pipeline slices on their own streams; one slice and odd slicings give the same jets
    assert 4096 >= 4 * model.pipeline_min_jets
    for chunks in (1, 3, 0, 2):   # one slice, odd slicing, direct mode (falls back to two slices: these tensors are not page-locked), two slices
        model.pipeline_chunks = chunks
        whole = model.simulate_dynamics(mk(), b, precision="fp32", jet_offset=0)
        assert torch.equal(out.continuous, whole.continuous) and torch.equal(out.discrete, whole.discrete)
    model.pipeline_chunks = 4
    part = model.simulate_dynamics(mk(slice(1000, 1100)), b, precision="fp32", jet_offset=1000)
    assert torch.equal(part.continuous, out.continuous[1000:1100]) and torch.equal(part.discrete, out.discrete[1000:1100])
    dead = b.source_mask == 0
    assert (out.discrete[dead] == 0).all() and (out.continuous[dead.expand(-1, -1, 3)] == 0).all()
    assert out.discrete.max() < 8 and torch.isfinite(out.continuous).all()
    assert (out.discrete != b.source_discrete).float().mean() > 0.1  # tokens did move
    dims, packed = ol.packed_model(model.cpu())
    sl = slice(2048, 2048 + 48)
    xo, ko = ol.generate(dims, packed, b.source_continuous[sl].numpy(), b.source_discrete[sl, :, 0].numpy(),
                         b.source_mask[sl, :, 0].numpy(), model.step_table(), seed=42, jet_offset=2048)
    assert np.array_equal(out.continuous[sl].numpy(), xo) and np.array_equal(out.discrete[sl, :, 0].numpy(), ko)


def test_validation_histograms_kernel_matches_host_statement():
    """csrc/histograms.cu == the torch statement of the same counts (sharding.ValidationHistograms on CPU)."""
    from multimodal_particles_b200 import sharding
    g = torch.Generator().manual_seed(4)
    b = jetclass_like_databatch(300, generator=g)
    x = (b.source_continuous * 2.5).contiguous()           # some values beyond [-5, 5): edge bins
    k = b.source_discrete[..., 0].to(torch.uint8).contiguous()
    m = b.source_mask[..., 0].to(torch.uint8).contiguous()
    hist = sharding.ValidationHistograms("cpu", vocab_size=8)
    want = hist.accumulate(x, k, m)
    got = hist.accumulate(x.to(DEV), k.to(DEV), m.to(DEV))
    assert got.is_cuda and torch.equal(got.cpu(), want)
    assert int(want[-129:].sum()) == 300 and int(want[:64].sum()) == int(m.sum())
