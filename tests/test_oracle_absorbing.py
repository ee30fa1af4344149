"""CPU: oracle restatement of the absorbing flow against the fixture of the unmodified reference
(tests/golden/make_golden_absorbing.py): trunk heads, transformer rate head, and the whole loop."""
import os

import numpy as np
import torch

import oracle_lib as ol


def test_state_dict_keys_and_param_count(golden_dir):
    z, cfg, model = ol.load_absorbing_golden(os.path.join(golden_dir, "absorbing.npz"))
    assert sum(p.numel() for p in model.parameters()) == 276682
    assert "generator.attn_blocks.1.proj_out.weight" in model.state_dict()


def test_heads_match_reference(golden_dir):
    z, cfg, model = ol.load_absorbing_golden(os.path.join(golden_dir, "absorbing.npz"))
    trunk, blob = ol.absorbing_trunk(model), model.generator.pack_head_weights().numpy()
    tab = model.step_table()
    g = model.generator
    for i in z["snap_steps"]:
        s = lambda name: z[f"snap{i}/{name}"]
        # host-side time tables equal the reference's
        np.testing.assert_allclose(g.time_bias(tab.t[i:i + 1]).numpy()[0], s("tbias")[0], rtol=1e-5, atol=1e-6)
        v, logits, hidden = ol.epic_forward(*trunk, s("x"), s("k")[..., 0], s("mask")[..., 0], tab.temb[i].numpy()[None], want_hidden=True)
        np.testing.assert_allclose(v, s("v"), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(logits, s("logits"), rtol=5e-5, atol=5e-5)
        a = ol.absorb_head(blob, g.encoder_output_dim_local, g.transformer_dim, g.n_heads, g.n_attn_blocks,
                           hidden, s("mask")[..., 0], s("tbias")[:1])
        np.testing.assert_allclose(a, s("a")[..., 0], rtol=2e-4, atol=2e-4)


def test_absorbing_loop_matches_reference(golden_dir):
    """birth -> Euler -> jump per step with oracle pieces == AbsorbingFlow.simulate_dynamics of the reference."""
    z, cfg, model = ol.load_absorbing_golden(os.path.join(golden_dir, "absorbing.npz"))
    trunk, blob = ol.absorbing_trunk(model), model.generator.pack_head_weights().numpy()
    tab = model.step_table()
    assert np.array_equal(tab.t.numpy(), z["t"]) and np.allclose(tab.sp.numpy(), z["sp"], rtol=1e-6)
    tb = model.generator.time_bias(tab.t).numpy()
    x, k, mask = z["x0"].copy(), z["k0"][..., 0].copy(), z["mask0"][..., 0].copy()
    for s in range(tab.n_steps):
        x, k, mask, _ = ol.absorbing_step(model, trunk, blob, x, k, mask, tab.temb[s].numpy()[None], tb[s][None],
                                          z["u_jump"][s], z["u_absorb"][s], tab.dt, float(tab.bc[s]), float(tab.cc[s]),
                                          float(tab.sp[s]))
        assert np.array_equal(mask, z["mask_traj"][s]), f"mask differs at step {s}"
    assert np.array_equal(mask, z["mask_final"][..., 0])
    same = (k == z["k_final"][..., 0]).all(-1)
    assert same.mean() >= 0.75
    np.testing.assert_allclose(x[same], z["x_final"][same], rtol=1e-4, atol=1e-4)
    assert mask.sum() > z["mask0"].sum()  # particles were born
