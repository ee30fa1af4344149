"""GPU: the jump rule inside the tensor-core generation kernels against the exact rule, on IDENTICAL logits and uniforms.

North star: "discrete token choices must be bit-exact given identical logits and uniforms".  That is proven for
``mmb_bridge_update`` and the fp32 generation kernel (tests/test_gpu_parity.py: bit-identical to the oracle, which reproduces the
reference's solver steps).  The bf16 / f16 generation kernels evaluate the same categorical with fast intrinsics (``__expf`` /
``ex2.approx``, approximate division, fused multiply-adds; ``telegraph_jump_fast*`` in csrc/mmb_device.cuh): their thresholds
differ from the exact ones in the last bits, so a draw that lands within rounding of a threshold can choose a neighbouring
token.  Here all three implementations get the same 5*10^7 draws at early / late / last-step coefficients; every disagreement is
listed with its distance to the nearest threshold (fp64) and must be such a draw.  SURVEY.md §7 expects O(10).
"""
import numpy as np
import pytest
import torch

from multimodal_particles_b200 import _native
from multimodal_particles_b200.steptable import build_step_table

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DRAWS_PER_REGIME = 10_000_000


def thresholds64(logits, k, dt, bc, cc):
    l = logits.astype(np.float64)
    q = np.exp(l - l.max(-1, keepdims=True))
    q /= q.sum(-1, keepdims=True)
    qk = np.take_along_axis(q, k[:, None].astype(np.int64), -1)
    lam = (1.0 + bc * q + cc * qk) * dt
    return np.cumsum(lam * np.exp(-lam.sum(-1, keepdims=True)), -1)


def test_fast_jump_rules_disagree_with_the_exact_rule_only_on_threshold_draws():
    tab = build_step_table(100, 1e-4, 8, 0.125, 16)
    g = torch.Generator(device=DEV).manual_seed(11)
    total = {"tc": 0, "mma": 0}
    worst = 0.0
    report = []
    for step, sigma in ((0, 1.0), (49, 4.0), (89, 2.0), (96, 4.0), (97, 1.0)):   # lambda(q=1) = 0.06 ... 7.9; the last step never moves
        P = DRAWS_PER_REGIME
        logits = torch.randn(P, 8, device=DEV, generator=g) * sigma
        k = torch.randint(0, 8, (P,), device=DEV, generator=g, dtype=torch.uint8)
        u = torch.rand(P, device=DEV, generator=g)
        dt, bc, cc = tab.dt, float(tab.bc[step]), float(tab.cc[step])
        exact, tc, mma = _native.jump_variants(logits, k, u, dt, bc, cc)
        moved = (exact != k).float().mean().item()
        for name, out in (("tc", tc), ("mma", mma)):
            idx = torch.nonzero(out != exact)[:, 0]
            total[name] += idx.numel()
            if idx.numel():
                li, ki, ui = logits[idx].cpu().numpy(), k[idx].cpu().numpy(), u[idx].cpu().numpy().astype(np.float64)
                c = thresholds64(li, ki, dt, bc, cc)
                dist = np.abs(ui[:, None] - c).min(-1)
                rel = dist / np.maximum(np.take_along_axis(c, np.abs(ui[:, None] - c).argmin(-1)[:, None], -1)[:, 0], 1e-30)
                worst = max(worst, float(rel.max()))
                report.append((step, name, int(idx.numel()), float(dist.max()), float(rel.max())))
                # a disagreement is a draw on a threshold: within 2e-6 relative (fast exp: 2 ulp; 8 accumulated terms)
                assert (rel < 2e-6).all(), f"step {step} {name}: draw {ui[rel.argmax()]} is {rel.max():.2e} (relative) from its threshold"
                # and the two answers are neighbours in the cumulative order (or "stay")
                a, b = out[idx].cpu().numpy().astype(int), exact[idx].cpu().numpy().astype(int)
                first = lambda tok: np.where(tok == ki, 8, tok)   # "stay" sorts after the last threshold
                assert (np.abs(first(a) - first(b)) <= 8).all()
        assert step >= 96 or moved > 0.02, (step, moved)
    print("jump variants: disagreements in 5e7 draws:", total, "worst relative distance to a threshold:", worst, report)
    assert total["tc"] <= 500 and total["mma"] <= 500, total   # ~1e-5 of the draws; measured O(10-100)


def test_jump_variants_equal_the_oracle_rule():
    """the `exact` output of the diagnostic kernel IS mmb_bridge_update's rule (bit-identical to the oracle)"""
    import oracle_lib as ol
    tab = build_step_table(100, 1e-4, 8, 0.125, 16)
    g = np.random.default_rng(3)
    P = 4096
    logits = (g.standard_normal((P, 8)) * 3).astype(np.float32)
    k = g.integers(0, 8, P).astype(np.uint8)
    u = g.random(P).astype(np.float32)
    step = 60
    dt, bc, cc = tab.dt, float(tab.bc[step]), float(tab.cc[step])
    _, ko, _ = ol.bridge_update(np.zeros((1, P, 3), np.float32), k[None], np.ones((1, P), np.uint8), np.zeros((1, P, 3), np.float32),
                                logits[None], u[None], dt, bc, cc)
    dev = lambda a: torch.from_numpy(a).to(DEV)
    exact, _, _ = _native.jump_variants(dev(logits), dev(k), dev(u), dt, bc, cc)
    assert np.array_equal(exact.cpu().numpy(), ko[0])
