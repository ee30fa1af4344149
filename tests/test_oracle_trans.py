"""The CPU oracle's trans-dimensional path against the fixture produced by the reference
(tests/golden/make_golden_trans.py): TransdimensionalEPiC.forward and JumpSampler.sample."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200.transdimensional import jump_schedule

GOLD = os.path.join(os.path.dirname(__file__), "golden", "trans.npz")


@pytest.fixture(scope="module")
def gold():
    z, cfg, model = ol.load_trans_golden(GOLD)
    return z, cfg, model, ol.trans_packed(model)


def test_blob_size_and_forward_rate(gold):
    z, cfg, model, packed = gold
    fr = model.forward_rate.as_c()
    ref = z["forward_rate"]
    assert fr.kind == 0 and abs(fr.scalar - ref[1]) < 1e-5 * ref[1] and fr.offset == np.float32(ref[2]) and fr.rate_cut_t == np.float32(ref[3])


def test_tokens_follow_the_batch_axis_softmax(gold):
    """structure.py:231-232: F.softmax without dim normalises a 3-D tensor over dim 0 — tokens depend on the batch."""
    z = gold[0]
    assert np.array_equal(ol.trans_tokens(z["fwd/onehot"]), z["fwd/tokens"])
    oh = torch.from_numpy(z["fwd/onehot"])
    assert not torch.equal(oh.argmax(-1), torch.from_numpy(z["fwd/tokens"]).long()), "fixture should exercise the quirk"


def test_forward_given_nearest(gold):
    z, cfg, model, packed = gold
    out = ol.trans_forward(packed, z["fwd/x"], z["fwd/onehot"], z["fwd/dims"], z["fwd/ts"], model.forward_rate.as_c(),
                           nearest=z["fwd/nearest"])
    np.testing.assert_allclose(out.d_xt, z["fwd/d_xt"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.x0_dim_logits, z["fwd/x0_dim_logits"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.near_atom_logits, z["fwd/near_atom_logits"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out.rate, z["fwd/rate"], rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(out.auto_mean, z["fwd/auto_mean"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out.auto_std, z["fwd/auto_std"], rtol=0, atol=5e-5)
    assert z["fwd/rate"][2] == 0.0 and (z["fwd/auto_mean"][2] == 0).all()   # dims == N: no birth possible


def test_forward_with_rate_use_x0_pred_false():
    """encoder.rate_use_x0_pred = False (transdimensional_model.py:185-188, 326-332): post_rate_proj has one output,
    rate = softplus(.) * forward_rate(t), x0_dim_logits = 0 — against the reference run of make_golden_trans_direct.py."""
    z, cfg, model = ol.load_trans_golden(os.path.join(os.path.dirname(GOLD), "trans_direct.npz"))
    assert cfg.encoder.rate_use_x0_pred is False and model.net.model.post_rate_proj.weight.shape[0] == 1
    packed = ol.trans_packed(model)
    assert packed[2].rate_direct == 1 and packed[2].max_particles == cfg.data.max_num_particles
    out = ol.trans_forward(packed, z["fwd/x"], z["fwd/onehot"], z["fwd/dims"], z["fwd/ts"], model.forward_rate.as_c(),
                           nearest=z["fwd/nearest"])
    np.testing.assert_allclose(out.d_xt, z["fwd/d_xt"], rtol=0, atol=2e-5)
    assert (out.x0_dim_logits == 0).all() and (z["fwd/x0_dim_logits"] == 0).all()
    np.testing.assert_allclose(out.near_atom_logits, z["fwd/near_atom_logits"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out.rate, z["fwd/rate"], rtol=2e-4, atol=1e-6)
    assert z["fwd/rate"].min() > 0 and z["fwd/rate"].max() > 20 * z["fwd/rate"].min()   # the fixture spans both sides of the rate cut
    np.testing.assert_allclose(out.auto_mean, z["fwd/auto_mean"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out.auto_std, z["fwd/auto_std"], rtol=0, atol=5e-5)


def test_forward_sampled_nearest(gold):
    z, cfg, model, packed = gold
    out = ol.trans_forward(packed, z["fwd/x"], z["fwd/onehot"], z["fwd/dims"], z["fwd/ts"], model.forward_rate.as_c(),
                           u_nearest=z["fwd/u_near"])
    assert np.array_equal(out.nearest, z["fwd/nearest_sampled"])
    np.testing.assert_allclose(out.auto_mean, z["fwd/auto_mean_sampled"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out.auto_std, z["fwd/auto_std_sampled"], rtol=0, atol=5e-5)


def test_schedule_equals_reference_loop(gold):
    z, cfg, model, packed = gold
    sched = jump_schedule(float(z["smp/dt"]), model.noise_schedule)
    assert sched.n_steps == len(z["smp/ts"]) and np.array_equal(sched.ts, z["smp/ts"])


def test_sampler_steps_one_by_one(gold):
    """every recorded step: oracle forward + update from the reference's state -> the reference's next state"""
    z, cfg, model, packed = gold
    fr = model.forward_rate.as_c()
    sched = jump_schedule(float(z["smp/dt"]), model.noise_schedule)
    n = sched.n_steps
    for i in range(n):
        x, oh, dims = z["smp/x_traj"][i], z["smp/oh_traj"][i], z["smp/dims_traj"][i]
        B, N, S = oh.shape
        ts = np.full(B, sched.ts[i], np.float32)
        f = ol.trans_forward(packed, x, oh, dims, ts, fr, u_nearest=z["smp/u_near"][i])
        assert np.array_equal(f.nearest, z["smp/nearest_traj"][i])
        np.testing.assert_allclose(f.rate, z["smp/rate_traj"][i], rtol=3e-4, atol=1e-5)
        v, lg = f.d_xt[:, :N * 3].reshape(B, N, 3), f.d_xt[:, N * 3:].reshape(B, N, S)
        x2, oh2, d2 = ol.trans_sampler_update(x, oh, dims, v, lg, f.rate, f.new_mean, f.new_std, sched.c_decay[i], sched.c_score[i],
                                              sched.c_noise[i], sched.inv_std[i], sched.jump_dt, z["smp/z_diff"][i],
                                              z["smp/u_jump"][i], z["smp/z_new"][i])
        if i + 1 < n:
            rx, ro, rd = z["smp/x_traj"][i + 1], z["smp/oh_traj"][i + 1], z["smp/dims_traj"][i + 1]
        else:
            rx, ro, rd = z["smp/x_final"], z["smp/oh_final"], z["smp/dims_final"]
        assert np.array_equal(d2, rd), f"step {i}"
        np.testing.assert_allclose(x2, rx, rtol=5e-6, atol=3e-5, err_msg=f"step {i}")
        np.testing.assert_allclose(oh2, ro, rtol=5e-6, atol=3e-5, err_msg=f"step {i}")


def test_sampler_whole_trajectory(gold):
    z, cfg, model, packed = gold
    B, N, S = z["smp/oh_final"].shape
    x, oh, dims = ol.trans_initial_state(z["smp/z_init"], N, S)
    np.testing.assert_allclose(oh, z["smp/oh_traj"][0], atol=0)
    np.testing.assert_allclose(x, z["smp/x_traj"][0], atol=1e-7)
    sched = jump_schedule(float(z["smp/dt"]), model.noise_schedule)
    x, oh, dims = ol.trans_sample(packed, x, oh, dims, sched, model.forward_rate.as_c(), z["smp/z_diff"], z["smp/u_near"],
                                  z["smp/u_jump"], z["smp/z_new"])
    assert np.array_equal(dims, z["smp/dims_final"])
    np.testing.assert_allclose(x, z["smp/x_final"], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(oh, z["smp/oh_final"], rtol=1e-4, atol=2e-3)


def test_sampler_c_time_grid(gold):
    """dt_schedule='C' + no_noise_final_step (sampler.py:79-88, 230): the reference's time grid and final state."""
    z, cfg, model, packed = gold
    B, N, S = z["smpC/oh_final"].shape
    sched = jump_schedule(float(z["smp/dt"]), model.noise_schedule, True, "C", 0.1, 0.04, 0.5)
    assert sched.n_steps == len(z["smpC/ts"]) and np.array_equal(sched.ts, z["smpC/ts"])
    assert sched.c_noise[-1] == 0.0 and np.all(sched.c_noise[:-1] > 0)
    n = sched.n_steps
    x, oh, dims = ol.trans_initial_state(z["smp/z_init"], N, S)
    x, oh, dims = ol.trans_sample(packed, x, oh, dims, sched, model.forward_rate.as_c(), z["smp/z_diff"][:n], z["smp/u_near"][:n],
                                  z["smp/u_jump"][:n], z["smp/z_new"][:n])
    assert np.array_equal(dims, z["smpC/dims_final"])
    np.testing.assert_allclose(x, z["smpC/x_final"], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(oh, z["smpC/oh_final"], rtol=1e-4, atol=2e-3)


def _corrector_schedule(z, model):
    import json
    kw = json.loads(str(z["smpL/kwargs"]))
    return jump_schedule(float(z["smpL/dt"]), model.noise_schedule, kw["no_noise_final_step"], corrector_steps=kw["corrector_steps"],
                         corrector_snr=kw["corrector_snr"], corrector_start_time=kw["corrector_start_time"],
                         corrector_finish_time=kw["corrector_finish_time"], do_jump_corrector=kw["do_jump_corrector"],
                         forward_rate=model.forward_rate)


def test_corrector_rows_follow_the_reference(gold):
    """Langevin corrector + jump corrector (sampler.py:258-312): every evaluation of the reference run, restarted from
    the reference's recorded state, lands on the reference's next state."""
    z, cfg, model, packed = gold
    sched = _corrector_schedule(z, model)
    assert sched.n_steps == len(z["smpL/ts"]) and np.array_equal(sched.ts, z["smpL/ts"])
    assert sched.kind.sum() == 24 and sched.c_noise[-1] == 0.0 and sched.kind[-1] == 1 and sched.c_noise[-3] > 0
    one = lambda i: SimpleNamespace(n_steps=1, kind=sched.kind[i:i + 1], ts=sched.ts[i:i + 1], c_decay=sched.c_decay[i:i + 1],
                                    c_score=sched.c_score[i:i + 1], c_noise=sched.c_noise[i:i + 1], inv_std=sched.inv_std[i:i + 1],
                                    death_prob=sched.death_prob[i:i + 1], jump_dt=sched.jump_dt, corrector_snr=sched.corrector_snr,
                                    jump_corrector=sched.jump_corrector)
    births = deaths = stale = 0
    for i in range(sched.n_steps):
        x, oh, dims = z["smpL/x_traj"][i], z["smpL/oh_traj"][i], z["smpL/dims_traj"][i]
        if sched.kind[i] == 0:
            mask_dims = dims   # the corrector rows reuse the mask of their predictor step (sampler.py:219)
        stale += int((mask_dims != dims).any())
        x2, oh2, d2 = ol.trans_sample(packed, x, oh, dims, one(i), model.forward_rate.as_c(), z["smpL/z_diff"][i:i + 1],
                                      z["smpL/u_near"][i:i + 1], z["smpL/u_jump"][i:i + 1], z["smpL/z_new"][i:i + 1],
                                      z["smpL/u_death"][i:i + 1], mask_dims)
        last = i + 1 == sched.n_steps
        rx, ro, rd = ((z["smpL/x_final"], z["smpL/oh_final"], z["smpL/dims_final"]) if last else
                      (z["smpL/x_traj"][i + 1], z["smpL/oh_traj"][i + 1], z["smpL/dims_traj"][i + 1]))
        assert np.array_equal(d2, rd), f"row {i}"
        births += int((rd > dims).sum()) if sched.kind[i] else 0
        deaths += int((rd < dims).sum())
        np.testing.assert_allclose(x2, rx, rtol=2e-5, atol=3e-5, err_msg=f"row {i}")
        np.testing.assert_allclose(oh2, ro, rtol=2e-5, atol=3e-5, err_msg=f"row {i}")
    assert births >= 1 and deaths >= 1 and stale >= 1   # both jumps of the corrector, and correctors with a stale mask


def test_corrector_whole_trajectory(gold):
    z, cfg, model, packed = gold
    B, N, S = z["smpL/oh_final"].shape
    sched = _corrector_schedule(z, model)
    x, oh, dims = ol.trans_initial_state(z["smpL/z_init"], N, S)
    x, oh, dims = ol.trans_sample(packed, x, oh, dims, sched, model.forward_rate.as_c(), z["smpL/z_diff"], z["smpL/u_near"],
                                  z["smpL/u_jump"], z["smpL/z_new"], z["smpL/u_death"])
    assert np.array_equal(dims, z["smpL/dims_final"])
    np.testing.assert_allclose(x, z["smpL/x_final"], rtol=1e-3, atol=5e-3)
    np.testing.assert_allclose(oh, z["smpL/oh_final"], rtol=1e-3, atol=5e-3)


def test_jump_sampler_constructor_accepts_the_optional_features_and_rejects_conditioning(gold):
    """host logic only: corrector steps, the jump corrector and the 'C' grid are built; conditioning is not (it needs the
    network's backward pass, sampler.py:133-135) and says so instead of silently sampling unconditionally"""
    from multimodal_particles_b200.transdimensional import JumpSampler
    z, cfg, model, packed = gold
    sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
    ok = JumpSampler(structure=model.structure, **dict(sk, corrector_steps=2, do_jump_corrector=True, dt_schedule="C",
                                                       no_noise_final_step=True))
    assert ok.corrector_steps == 2 and ok.do_jump_corrector and abs(float(ok.get_dt(torch.tensor([0.9]))) - sk["dt_schedule_h"]) < 1e-9
    for bad in (dict(do_conditioning=True), dict(dt_schedule="quadratic"), dict(sample_near_atom=False)):
        with pytest.raises(NotImplementedError):
            JumpSampler(structure=model.structure, **dict(sk, **bad))
