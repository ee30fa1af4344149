"""Source-state sampler: oracle against the reference's distributions (CPU), kernel against the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200.source import multiplicity_cdf

GOLD = os.path.join(os.path.dirname(__file__), "golden", "source.npz")


def expected_token_probs(cat_probs):
    p = np.asarray(cat_probs, np.float64)
    return np.array([p[0], p[1], p[2] / 2, p[2] / 2, p[3] / 2, p[3] / 2, p[4] / 2, p[4] / 2])


def check_distribution(z, x, k, mask, n_sigma=4.5):
    B, N = k.shape
    live = mask.astype(bool)
    assert (mask[:, 1:] <= mask[:, :-1]).all()                       # prefix masks
    assert (k[~live] == 0).all() and (x[~live] == 0).all()           # masked slots are zeroed (particles.py:67-69)
    n = live.sum()
    freq = np.bincount(k[live], minlength=8) / n
    want = expected_token_probs(z["cat_probs"])
    assert (np.abs(freq - want) <= n_sigma * np.sqrt(want * (1 - want) / n) + 1e-9).all(), (freq, want)
    assert (np.abs(z["token_freq"] - want) <= n_sigma * np.sqrt(want / 600000)).all()       # the reference agrees with the same law
    mult = mask.sum(1)
    cdf = multiplicity_cdf(z["hist"], N)
    pm = np.diff(np.concatenate([[0.0], cdf.astype(np.float64)]))
    fm = np.bincount(mult, minlength=N + 1) / B
    assert (np.abs(fm - pm) <= n_sigma * np.sqrt(pm * (1 - pm) / B) + 1e-9).all()
    assert (np.abs(z["mult_freq"] - pm) <= n_sigma * np.sqrt(pm * (1 - pm) / 20000) + 1e-9).all()
    xs = x[live] / float(z["scale"])
    assert np.abs(xs.mean(0)).max() < 5 / np.sqrt(n) and np.abs(xs.std(0) - 1).max() < 5 / np.sqrt(n)
    assert abs((xs ** 4).mean() - 3.0) < 0.05


def test_oracle_source_follows_reference_distributions():
    z = np.load(GOLD)
    x, k, mask = ol.sample_source(20000, 30, float(z["scale"]), z["cat_probs"], multiplicity_cdf(z["hist"], 30), 7, 0)
    check_distribution(z, x, k, mask)


@pytest.mark.gpu
def test_kernel_matches_oracle_and_is_shard_invariant():
    from multimodal_particles_b200.source import sample_source_state
    z = np.load(GOLD)
    kw = dict(max_num_particles=30, target_multiplicity=z["hist"], scale=float(z["scale"]), cat_probs=z["cat_probs"], seed=7)
    x, k, mask = sample_source_state(20000, compact=True, **kw)
    wx, wk, wm = ol.sample_source(20000, 30, float(z["scale"]), z["cat_probs"], multiplicity_cdf(z["hist"], 30), 7, 0)
    assert np.array_equal(k.cpu().numpy(), wk) and np.array_equal(mask.cpu().numpy(), wm)      # integer side: bit-exact
    np.testing.assert_allclose(x.cpu().numpy(), wx, rtol=0, atol=2e-4 * float(z["scale"]) * 5)   # fast log / sincos in Box-Muller
    check_distribution(z, x.cpu().numpy(), k.cpu().numpy(), mask.cpu().numpy())
    x2, k2, m2 = sample_source_state(500, compact=True, jet_offset=1000, **kw)
    assert torch.equal(k2, k[1000:1500]) and torch.equal(m2, mask[1000:1500]) and torch.equal(x2, x[1000:1500])
    st = sample_source_state(64, max_num_particles=128, seed=1)      # reference layout, full jets
    assert st.continuous.shape == (64, 128, 3) and st.discrete.dtype == torch.int64 and st.absorbing.shape == (64, 128, 1)
    assert int(st.absorbing.sum()) == 64 * 128
