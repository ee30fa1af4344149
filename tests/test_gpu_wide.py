"""The 128-wide tcgen05 trunk (csrc/epic_wide_tc.cu; EPiCNetwork's class-default widths, epic.py:99-101) through the C ABI:
against the reference fixture at those widths, against the fp32 kernel (itself bit-exact against the oracle at any width), and as
the network of a whole generation.  bf16 operands, fp32 accumulation: heads within 2e-2 of the largest head (measured 3-6e-3)."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200 import HybridState, MultiModalBridgeMatching
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.epic import as_u8

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL = 2e-2


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(DEV, dtype) if dtype else t.to(DEV)


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def wide_model(seed=0, L=6, G=10, skip=True, head=True, ctx=0, S=8):
    cfg = MultimodalBridgeMatchingConfig()
    e = cfg.encoder
    e.dim_hidden_local, e.num_blocks, e.dim_hidden_glob, e.skip_connection, e.add_discrete_head = 128, L, G, skip, head
    cfg.data.dim_context_continuous, cfg.data.vocab_size_features = ctx, S
    cfg.bridge.num_timesteps = 30
    torch.manual_seed(seed)
    model = MultiModalBridgeMatching(cfg).to(DEV)
    with torch.no_grad():   # sharpen the random-init heads so that tokens compete (as in the fixtures)
        if head:
            model.encoder.fc_layer[2].weight.mul_(6.0)
        model.encoder.epic.epic.output_layer.weight_g.mul_(3.0)
    return cfg, model


def test_fixture_heads_at_class_default_widths(golden_dir):
    """Reference heads (tests/golden/mbm_wide.npz) vs the tcgen05 trunk; the fp32 kernel on the same states is bit-exact
    against the oracle."""
    z, cfg, model = ol.load_mbm_golden(os.path.join(golden_dir, "mbm_wide.npz"))
    dims, packed = ol.packed_model(model)
    assert (dims.dim_hidden_local, dims.num_blocks, dims.dim_hidden_glob) == (128, 6, 10)
    native = model.to(DEV).encoder.native_model(torch.device(DEV))
    assert native.generate_precision(128) == "bf16"   # "auto": the warp-MMA engine is H = 16 only, the wide trunk takes it
    for i in z["snap_steps"]:
        x, k, mask, temb = z[f"snap{i}/x"], z[f"snap{i}/k"][..., 0], z["mask"][..., 0], z["temb"][i][None]
        vw, lw = ol.epic_forward(dims, packed, x, k, mask, temb)
        v32, l32 = native.forward(dev(x), dev(k), dev(mask), dev(temb), precision="fp32")
        assert np.array_equal(v32.cpu().numpy().view(np.uint32), vw.view(np.uint32))
        assert np.array_equal(l32.cpu().numpy().view(np.uint32), lw.view(np.uint32))
        v, lg = native.forward(dev(x), dev(k), dev(mask), dev(temb), precision="bf16")
        assert rel(v.cpu(), torch.from_numpy(z[f"snap{i}/v"])) < REL
        assert rel(lg.cpu(), torch.from_numpy(z[f"snap{i}/logits"])) < REL


@pytest.mark.parametrize("kw", [dict(), dict(L=2, G=16, skip=False, head=False), dict(L=3, G=7, ctx=5), dict(L=1, G=32, S=4), dict(L=8, G=1)])
@pytest.mark.parametrize("B,N", [(37, 128), (2, 50), (1, 1)])
def test_forward_tracks_fp32(kw, B, N):
    """Odd and even jet counts (a CTA carries two jets), partial tiles, every supported width; dead rows and hidden states."""
    cfg, model = wide_model(**kw)
    native = model.encoder.native_model(torch.device(DEV))
    b = jetclass_like_databatch(B, N, generator=torch.Generator().manual_seed(5 + B))
    x, k, m = b.source_continuous.to(DEV), as_u8((b.source_discrete % cfg.data.vocab_size_features).to(DEV)), as_u8(b.source_mask.to(DEV))
    x = x + (1 - m[..., None].float()) * 7.0     # junk on padding: nothing may depend on it
    temb = torch.randn(B, cfg.encoder.dim_emb_time + kw.get("ctx", 0), device=DEV)   # per-jet times (and contexts)
    v0, l0, h0 = native.forward(x, k, m, temb, want_hidden=True, precision="fp32")
    v1, l1, h1 = native.forward(x, k, m, temb, want_hidden=True, precision="bf16")
    assert torch.isfinite(v1).all() and torch.isfinite(l1).all() and torch.isfinite(h1).all()
    assert rel(v1, v0) < REL and rel(l1, l0) < REL and rel(h1, h0) < REL
    dead = m == 0
    assert (v1[dead] == 0).all() and (h1[dead] == 0).all()
    # head(0): the reference applies the head to the masked logits (mbm.py:105-111); same value up to the SELU's exp intrinsic
    assert torch.allclose(l1[dead], l0[dead], rtol=1e-5, atol=1e-5)
    v2, l2 = native.forward(x, k, m, temb, precision="bf16")
    assert torch.equal(v1, v2) and torch.equal(l1, l2)   # deterministic
    if B > 2:   # a jet's result does not depend on its partner in the CTA or its position in the batch
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(DEV)
        v3, l3 = native.forward(x[perm].contiguous(), k[perm].contiguous(), m[perm].contiguous(), temb[perm].contiguous(), precision="bf16")
        assert torch.equal(v3, v1[perm]) and torch.equal(l3, l1[perm])


@pytest.mark.parametrize("mults", [[40, 50, 64], [1, 2, 3, 4, 5], [96, 70, 10], [65, 80, 90, 5, 6, 7, 8, 9], [128], [128, 1], [33], [32, 32, 32],
                                   [97, 128, 100, 3], [64, 64, 64, 64, 1]])
def test_tile_compositions(mults):
    """Every way the pre-pass composes tiles — [4] | [3,1] | [3] | [2,2] | [2,1] | [2] | [1,1] | [1] — with odd counts in each class."""
    cfg, model = wide_model(L=2, seed=7)
    native = model.encoder.native_model(torch.device(DEV))
    B, N = len(mults), 128
    g = torch.Generator().manual_seed(sum(mults))
    m = (torch.arange(N)[None] < torch.tensor(mults)[:, None]).to(torch.uint8)
    x = (torch.randn(B, N, 3, generator=g) * m[..., None]).to(DEV)
    k = (torch.randint(0, 8, (B, N), generator=g, dtype=torch.uint8) * m).to(DEV)
    m = m.to(DEV)
    temb = torch.randn(B, cfg.encoder.dim_emb_time, generator=g).to(DEV)
    v0, l0, h0 = native.forward(x, k, m, temb, want_hidden=True, precision="fp32")
    v1, l1, h1 = native.forward(x, k, m, temb, want_hidden=True, precision="bf16")
    for j in range(B):   # per jet: nobody is skipped, nobody gets a neighbour's result
        assert rel(v1[j], v0[j]) < REL and rel(l1[j], l0[j]) < REL and rel(h1[j], h0[j]) < REL, (mults, j)


def test_full_size_batch_and_every_multiplicity():
    """4096 jets (the C2 batch) with multiplicities 1..128 all present, incl. full jets ([4] tiles) and one-particle jets
    ([1,1] tiles): every head against the fp32 kernel; jets ordered by multiplicity and shuffled give the same bits."""
    cfg, model = wide_model(seed=5)
    native = model.encoder.native_model(torch.device(DEV))
    B, N = 4096, 128
    g = torch.Generator().manual_seed(3)
    mult = torch.cat([torch.arange(1, N + 1), torch.randint(1, N + 1, (B - N,), generator=g)])
    m = (torch.arange(N)[None] < mult[:, None]).to(torch.uint8)
    m = torch.gather(m, 1, torch.rand(B, N, generator=g).argsort(1))          # live slots anywhere, not a prefix
    x = (torch.randn(B, N, 3, generator=g) * m[..., None]).to(DEV)
    k = (torch.randint(0, 8, (B, N), generator=g, dtype=torch.uint8) * m).to(DEV)
    m = m.to(DEV)
    temb = model.step_table().temb[7:8].to(DEV)
    v0, l0 = native.forward(x, k, m, temb, precision="fp32")
    v1, l1 = native.forward(x, k, m, temb, precision="bf16")
    assert rel(v1, v0) < REL and rel(l1, l0) < REL
    per_jet = (v1 - v0).abs().amax((1, 2)) / v0.abs().amax()
    assert per_jet.max().item() < REL, int(per_jet.argmax())
    perm = torch.randperm(B, generator=g).to(DEV)
    v2, l2 = native.forward(x[perm].contiguous(), k[perm].contiguous(), m[perm].contiguous(), temb, precision="bf16")
    assert torch.equal(v2, v1[perm]) and torch.equal(l2, l1[perm])


def test_empty_jet_is_nan_like_the_reference_and_leaves_its_partner_alone():
    cfg, model = wide_model(L=2)
    native = model.encoder.native_model(torch.device(DEV))
    b = jetclass_like_databatch(4, 128, generator=torch.Generator().manual_seed(9))
    x, k, m = b.source_continuous.to(DEV), as_u8(b.source_discrete.to(DEV)), as_u8(b.source_mask.to(DEV))
    temb = torch.randn(1, cfg.encoder.dim_emb_time, device=DEV)
    v_ref, l_ref = native.forward(x, k, m, temb, precision="bf16")
    m2 = m.clone(); m2[1] = 0
    v, lg = native.forward(x, k, m2, temb, precision="bf16")
    assert torch.equal(v[0], v_ref[0]) and torch.equal(v[2:], v_ref[2:]) and torch.equal(lg[0], l_ref[0])
    v32, l32 = native.forward(x, k, m2, temb, precision="fp32")
    # the mean pool of a jet without particles is 0 / 0 (epic.py:141) and NaN * mask stays NaN: every head of the jet is NaN
    assert torch.isnan(v[1]).all() and torch.isnan(v32[1]).all() and torch.isnan(lg[1]).all() and torch.isnan(l32[1]).all()


def test_generation_with_the_wide_trunk():
    """mmb_generate(bf16) on a class-default-width model = per step (tcgen05 trunk, fused update kernel, Philox draws of that
    step): same draws as the fp32 loop, so tokens agree except near thresholds; host and device entry points agree bit for bit."""
    cfg, model = wide_model(seed=3)
    B = 256
    b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(11))
    mk = lambda where: HybridState(None, *[t.clone().to(where) for t in (b.source_continuous, b.source_discrete, b.source_mask)])
    model.seed = 4
    ref = model.simulate_dynamics(mk(DEV), b, precision="fp32", jet_offset=50)
    got = model.simulate_dynamics(mk(DEV), b, precision="bf16", jet_offset=50)
    live = b.source_mask[..., 0].bool()
    moved = (ref.discrete != b.source_discrete)[live].float().mean().item()
    agree = (ref.discrete == got.discrete)[live].float().mean().item()
    dx = (ref.continuous - got.continuous).abs()[live].mean().item()
    travel = (ref.continuous - b.source_continuous).abs()[live].mean().item()   # how far the features move over the bridge
    print(f"wide generation, bf16 vs fp32: tokens moved {moved:.3f}, agreement {agree:.4f}, mean |dx| {dx:.5f}, mean travel {travel:.3f}")
    assert moved > 0.2 and agree >= 0.95 and dx <= 0.02 * travel
    assert (got.discrete[~live] == 0).all() and (got.continuous[~live] == 0).all()
    auto = model.simulate_dynamics(mk(DEV), b, jet_offset=50)          # precision "auto" -> bf16 for this model
    assert torch.equal(auto.continuous, got.continuous) and torch.equal(auto.discrete, got.discrete)
    host = model.simulate_dynamics(mk("cpu"), b, jet_offset=50)        # host tensors in: mmb_generate_host (sliced mode)
    assert torch.equal(host.continuous, got.continuous) and torch.equal(host.discrete, got.discrete)
    u = torch.rand(cfg.bridge.num_timesteps - 1, B, 128, generator=torch.Generator().manual_seed(2))
    a = model.simulate_dynamics(mk(DEV), b, uniforms=u, precision="bf16")
    c = model.simulate_dynamics(mk(DEV), b, uniforms=u, precision="fp32")
    assert (a.discrete == c.discrete)[live].float().mean().item() >= 0.95
