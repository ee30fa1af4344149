"""GPU: the warp-MMA generation engine (csrc/epic_mma.cu, precision "f16") and the C2-size parity of both tensor-core engines.

Stated tolerances against the fp32 path (bit-identical to the oracle), same injected uniforms:
* golden cases: >= 90 % identical tokens, mean |dx| <= 0.05 (as tests/test_gpu_tc.py states for bf16; measured f16: 99.6-100 %);
* C2 full size (4096 jets, JetClass-like masks, 99 steps, heads sharpened so that 87 % of the tokens move): bf16 >= 95 %
  identical tokens, f16 >= 98 %, mean |dx| <= 0.02 / 0.01 (measured values are printed), and the 1-D Wasserstein distances of features / jet sums / token frequencies to
  the fp32 twin below the distance between two independent fp32 samples.
Empty jets (no live particle): NaN features and zero tokens in EVERY precision, as the reference's 0/0 mean pool gives
(epic.py:141, bridges.py:42); jets that share a tile / CTA / call with an empty jet are bit-identical to their solo runs.
"""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200 import HybridState, MultiModalBridgeMatching, _native
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.epic import as_u8
from test_gpu_tc import golden_model, w1

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sharp_model(num_timesteps=100, seed=0):
    cfg = MultimodalBridgeMatchingConfig()
    cfg.bridge.num_timesteps = num_timesteps
    torch.manual_seed(seed)
    model = MultiModalBridgeMatching(cfg).to(DEV)
    with torch.no_grad():  # sharpen the random-init heads so tokens and features actually move
        model.encoder.fc_layer[2].weight.mul_(6.0)
        model.encoder.epic.epic.output_layer.weight_g.mul_(3.0)
    return cfg, model


def state_of(b, sl=slice(None)):
    return HybridState(None, b.source_continuous[sl].clone(), b.source_discrete[sl].clone(), b.source_mask[sl].clone())


@pytest.mark.parametrize("case", ["mbm_c1", "mbm_n128"])
def test_generation_f16_tracks_fp32(case, golden_dir):
    z, cfg, model = golden_model(golden_dir, case)
    assert model.encoder.native_model(torch.device(DEV)).generate_precision(z["x0"].shape[1]) == "f16"   # what "auto" picks
    mk = lambda: HybridState(None, torch.from_numpy(z["x0"]), torch.from_numpy(z["k0"]).long(), torch.from_numpy(z["mask"]).long())
    u = torch.from_numpy(z["u_jump"])
    a = model.simulate_dynamics(mk(), None, uniforms=u, precision="fp32")
    b = model.simulate_dynamics(mk(), None, uniforms=u, precision="f16")
    live = torch.from_numpy(z["mask"]).bool()
    assert (a.discrete == b.discrete)[live].float().mean() >= 0.90
    assert (a.continuous - b.continuous).abs().mean() <= 0.05
    assert (b.discrete[~live] == 0).all() and (b.continuous[(~live).expand(-1, -1, 3)] == 0).all()
    # in-kernel Philox == the documented (seed, jet, step, particle) -> uniform map, bit for bit
    B, N = z["x0"].shape[:2]
    model.seed = 77
    up = _native.philox_uniforms(77, 1000, model.step_table().n_steps, B, N, torch.device(DEV))
    c = model.simulate_dynamics(mk(), None, precision="f16", jet_offset=1000)
    d = model.simulate_dynamics(mk(), None, uniforms=up, precision="f16")
    assert torch.equal(c.discrete, d.discrete) and torch.equal(c.continuous, d.continuous)


def test_f16_jets_do_not_depend_on_the_call_they_are_generated_in():
    """The f16 engine bins jets by width and lets warps claim them dynamically; a jet wider than 64 particles spans two warps
    that exchange partial pooling sums.  Whatever the call contains, every jet comes out bit-identical."""
    cfg, model = sharp_model(20, seed=3)
    b = jetclass_like_databatch(515, generator=torch.Generator().manual_seed(5))
    mult = b.source_mask[..., 0].sum(1)
    assert (mult <= 16).sum() > 5 and ((mult > 32) & (mult <= 48)).sum() > 50 and (mult > 64).sum() > 20 and (mult > 80).sum() > 3
    model.seed = 11
    whole = model.simulate_dynamics(state_of(b), None, precision="f16", jet_offset=0)
    lo = 0
    for size in [4, 1, 3, 17] * 40:
        if lo >= 515:
            break
        part = model.simulate_dynamics(state_of(b, slice(lo, lo + size)), None, precision="f16", jet_offset=lo)
        assert torch.equal(part.continuous, whole.continuous[lo:lo + size]) and torch.equal(part.discrete, whole.discrete[lo:lo + size]), (lo, size)
        lo += size
    dead = b.source_mask == 0
    assert (whole.discrete[dead] == 0).all() and (whole.continuous[dead.expand(-1, -1, 3)] == 0).all()
    assert (whole.discrete != b.source_discrete).float().mean() > 0.05
    # wide jets: N = 200 particles per jet (four m-tiles x up to four warps), random (non-prefix) masks
    cfg2 = MultimodalBridgeMatchingConfig()
    cfg2.bridge.num_timesteps, cfg2.data.max_num_particles = 12, 200
    torch.manual_seed(4)
    wide = MultiModalBridgeMatching(cfg2).to(DEV)
    g = torch.Generator().manual_seed(9)
    B, N = 37, 200
    mask = (torch.rand(B, N, 1, generator=g) < torch.rand(B, 1, 1, generator=g)).long()
    mask[:, 0] = 1
    x0, k0 = torch.randn(B, N, 3, generator=g) * mask, torch.randint(0, 8, (B, N, 1), generator=g) * mask
    u = torch.rand(11, B, N, generator=g)
    mk = lambda sl=slice(None): HybridState(None, x0[sl].clone(), k0[sl].clone(), mask[sl].clone())
    f32 = wide.simulate_dynamics(mk(), None, uniforms=u, precision="fp32")
    f16 = wide.simulate_dynamics(mk(), None, uniforms=u, precision="f16")
    assert (f32.discrete == f16.discrete)[mask.bool()].float().mean() >= 0.97
    assert (f32.continuous - f16.continuous).abs().max() <= 0.02
    solo = wide.simulate_dynamics(mk(slice(5, 6)), None, uniforms=u[:, 5:6], precision="f16")
    assert torch.equal(solo.continuous, f16.continuous[5:6]) and torch.equal(solo.discrete, f16.discrete[5:6])


def test_c2_full_size_tensor_core_engines_track_fp32():
    """BASELINE config 2 at full size: both tensor-core engines against the fp32 path on the same 4096 jets and uniforms."""
    cfg, model = sharp_model()
    B, N = 4096, 128
    ba = jetclass_like_databatch(B, N, generator=torch.Generator().manual_seed(1234))
    bb = jetclass_like_databatch(B, N, generator=torch.Generator().manual_seed(4321))
    table = model.step_table()
    u = _native.philox_uniforms(5, 0, table.n_steps, B, N, torch.device(DEV))
    ub = _native.philox_uniforms(6, 0, table.n_steps, B, N, torch.device(DEV))
    run = lambda b, prec, uu: model.simulate_dynamics(state_of(b), b, uniforms=uu, precision=prec)
    fa, fb = run(ba, "fp32", u), run(bb, "fp32", ub)
    la, lb = ba.source_mask[..., 0].bool(), bb.source_mask[..., 0].bool()
    assert (fa.discrete != ba.source_discrete)[la].float().mean() > 0.3
    freq = lambda s, live: np.bincount(s.discrete[..., 0][live].numpy(), minlength=8) / int(live.sum())
    for prec, min_agree, max_dx in (("bf16", 0.95, 0.02), ("f16", 0.98, 0.01)):
        t = run(ba, prec, u)
        agree = (t.discrete == fa.discrete)[la].float().mean().item()
        dx = (t.continuous - fa.continuous).abs()[la].mean().item()
        print(f"C2 full size, {prec} vs fp32: token agreement {agree:.4f}, mean |dx| {dx:.5f}")
        assert agree >= min_agree, (prec, agree)
        assert dx <= max_dx, (prec, dx)
        assert (t.discrete[~la] == 0).all() and (t.continuous[(~la)[..., None].expand(-1, -1, 3)] == 0).all()
        for c in range(3):
            spread = w1(fa.continuous[..., c][la], fb.continuous[..., c][lb])
            assert w1(t.continuous[..., c][la], fa.continuous[..., c][la]) <= spread, (prec, c)
            spread = w1(fa.continuous[..., c].sum(1), fb.continuous[..., c].sum(1))
            assert w1(t.continuous[..., c].sum(1), fa.continuous[..., c].sum(1)) <= spread, (prec, "jet sum", c)
        assert np.abs(freq(t, la) - freq(fa, la)).sum() <= np.abs(freq(fa, la) - freq(fb, lb)).sum()


@pytest.mark.parametrize("precision", ["fp32", "bf16", "f16"])
def test_empty_jets_are_nan_like_the_reference_and_leave_their_neighbours_alone(precision):
    cfg, model = sharp_model(20, seed=8)
    b = jetclass_like_databatch(64, generator=torch.Generator().manual_seed(21))
    mask = b.source_mask.clone()
    empty = [0, 5, 6, 33, 63]
    mask[empty] = 0
    x0 = b.source_continuous * mask + (1 - mask) * 7.0     # junk on padding: the result must not depend on it
    k0 = b.source_discrete * mask
    mk = lambda sl=slice(None): HybridState(None, x0[sl].clone(), k0[sl].clone(), mask[sl].clone())
    model.seed = 3
    out = model.simulate_dynamics(mk(), None, precision=precision, jet_offset=100)
    for j in empty:   # reference: mean pool 0/0 -> NaN velocity -> (x + dt NaN) * 0 = NaN for every particle; k * mask = 0
        assert torch.isnan(out.continuous[j]).all(), (precision, j)
        assert (out.discrete[j] == 0).all()
    live_jets = [j for j in range(64) if j not in empty]
    assert torch.isfinite(out.continuous[live_jets]).all()
    dead = (mask == 0)
    dead[empty] = False
    assert (out.continuous[dead.expand(-1, -1, 3)] == 0).all() and (out.discrete[dead] == 0).all()
    # the neighbours of an empty jet (same tile / CTA / call) equal their solo runs bit for bit
    for j in (1, 4, 7, 32, 34, 62):
        solo = model.simulate_dynamics(mk(slice(j, j + 1)), None, precision=precision, jet_offset=100 + j)
        assert torch.equal(solo.continuous[0], out.continuous[j]) and torch.equal(solo.discrete[0], out.discrete[j]), (precision, j)
    # an empty jet alone
    alone = model.simulate_dynamics(mk(slice(5, 6)), None, precision=precision, jet_offset=105)
    assert torch.isnan(alone.continuous).all() and (alone.discrete == 0).all()


def test_host_pipeline_modes_give_identical_jets():
    """simulate_dynamics with host tensors: direct mode (pinned buffers, the f16 kernel reads / writes them itself), sliced
    modes and the device-resident call produce the same jets bit for bit; the token-range assertion fires in every mode."""
    cfg, model = sharp_model(30, seed=5)
    b = jetclass_like_databatch(3000, generator=torch.Generator().manual_seed(31))
    model.seed, model.pipeline_min_jets = 9, 1
    pin = lambda t: t.clone().pin_memory()
    mk = lambda pinned: HybridState(None, *(pin(t) if pinned else t.clone() for t in (b.source_continuous, b.source_discrete, b.source_mask)))
    dev_state = HybridState(None, b.source_continuous.to(DEV), b.source_discrete.to(DEV), b.source_mask.to(DEV))
    ref = model.simulate_dynamics(dev_state, None, precision="f16", jet_offset=50, return_device=True)
    for chunks, pinned in ((0, True), (0, False), (1, True), (3, True), (2, False)):
        model.pipeline_chunks = chunks
        out = model.simulate_dynamics(mk(pinned), None, precision="f16", jet_offset=50)
        assert out.continuous.device.type == "cpu" and out.discrete.dtype == torch.int64 and out.discrete.shape == (3000, 128, 1)
        assert torch.equal(out.continuous, ref.continuous.cpu()) and torch.equal(out.discrete, ref.discrete.cpu()), (chunks, pinned)
    # empty jets through the direct mode: NaN features, zero tokens (written by the binning prologue)
    st = mk(True)
    st.absorbing[7] = 0
    model.pipeline_chunks = 0
    out = model.simulate_dynamics(st, None, precision="f16", jet_offset=50)
    assert torch.isnan(out.continuous[7]).all() and (out.discrete[7] == 0).all() and torch.equal(out.continuous[8], ref.continuous[8].cpu())
    for chunks in (0, 2):
        model.pipeline_chunks = chunks
        bad = mk(True)
        bad.discrete[17, 3, 0] = 8
        with pytest.raises(AssertionError):
            model.simulate_dynamics(bad, None, precision="f16", jet_offset=50)
        bad = mk(True)
        bad.discrete[17, 120, 0] = -1      # a dead particle's token: the reference asserts on every entry
        with pytest.raises(AssertionError):
            model.simulate_dynamics(bad, None, precision="f16", jet_offset=50)
