"""GPU: the trans-dimensional jump diffusion (BASELINE configs 3/4) through the C ABI,
against the fp32 oracle and the fixture produced by the reference (tests/golden/make_golden_trans.py).

Stated tolerances.  The two transformer stacks run ~16 chained bf16 GEMMs, three GroupNorms and a softmax per block:
per-particle / per-jet head outputs within 3 % of the largest reference magnitude (as for the absorbing head).
The trunk in fp32 is the oracle's arithmetic; its time embedding is evaluated with CUDA sinf/cosf instead of libm, so
D_xt agrees to 2e-5 of its scale rather than bit for bit.  The birth rate is a smooth function of the x0-dimension
logits: 5 % relative.  Discrete decisions (tokens, births, nearest particle) are bit-exact given identical inputs:
the fused sampler update is checked against the oracle with the oracle's own network outputs fed to both sides.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import oracle_lib as ol
from multimodal_particles_b200 import _native
from multimodal_particles_b200.config_classes.transdimensional_unconditional_config import TransdimensionalEpicConfig
from multimodal_particles_b200.transdimensional import (JumpSampler, StructuredDataBatch, TransdimensionalJumpDiffusion,
                                                        jump_schedule)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def close(got, want, frac, floor=1e-6):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    assert np.isfinite(got).all()
    err, scale = np.abs(got - want).max(), max(np.abs(want).max(), floor)
    assert err <= frac * scale, (err, scale)


@pytest.fixture(scope="module")
def fixture(golden_dir):
    z, cfg, model = ol.load_trans_golden(os.path.join(golden_dir, "trans.npz"))
    packed = ol.trans_packed(model)
    return z, cfg, model.to(DEV), packed


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_given_nearest_matches_reference(fixture, precision):
    z, cfg, model, packed = fixture
    net = model.net
    net.model.precision = precision
    st = model.make_batch(to_dev(z["fwd/x"]), to_dev(z["fwd/onehot"]), to_dev(z["fwd/dims"]))
    B, N, S = z["fwd/onehot"].shape
    D, rate, (am, asd), x0l, nal = net(st, to_dev(z["fwd/ts"]), forward_rate=model.forward_rate, predict="eps",
                                       nearest_atom=to_dev(z["fwd/nearest"]))
    assert D.shape == (B, N * (3 + S)) and rate.shape == (B, 1) and am.shape == D.shape and x0l.shape == (B, N) and nal.shape == (B, N)
    close(D, z["fwd/d_xt"], 2e-5 if precision == "fp32" else 0.02)
    close(x0l, z["fwd/x0_dim_logits"], 0.03)
    close(nal, z["fwd/near_atom_logits"], 0.03)
    np.testing.assert_allclose(rate.view(-1).cpu().numpy(), z["fwd/rate"], rtol=0.05, atol=1e-4)
    close(am, z["fwd/auto_mean"], 0.03)
    close(asd, z["fwd/auto_std"], 0.03)
    # only the slot a birth fills is non-zero (structure.py:175-184); none when the jet is full
    assert ((am.cpu().numpy() != 0) == (z["fwd/auto_mean"] != 0)).all()
    assert torch.equal(net.model.last_nearest_atom.cpu(), torch.from_numpy(z["fwd/nearest"]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_with_rate_use_x0_pred_false(golden_dir, precision):
    """encoder.rate_use_x0_pred = False: one rate logit, rate = softplus(.) * forward_rate(t), zero x0_dim_logits
    (transdimensional_model.py:185-188, 326-332), against the reference run and the oracle; the sampler runs on such a model."""
    z, cfg, model = ol.load_trans_golden(os.path.join(golden_dir, "trans_direct.npz"))
    packed = ol.trans_packed(model)
    model = model.to(DEV)
    net = model.net
    net.model.precision = precision
    st = model.make_batch(to_dev(z["fwd/x"]), to_dev(z["fwd/onehot"]), to_dev(z["fwd/dims"]))
    B, N, S = z["fwd/onehot"].shape
    D, rate, (am, asd), x0l, nal = net(st, to_dev(z["fwd/ts"]), forward_rate=model.forward_rate, predict="eps",
                                       nearest_atom=to_dev(z["fwd/nearest"]))
    assert rate.shape == (B, 1) and x0l.shape == (B, N) and (x0l == 0).all()
    close(D, z["fwd/d_xt"], 2e-5 if precision == "fp32" else 0.02)
    close(nal, z["fwd/near_atom_logits"], 0.03)
    # the rate logit comes out of the bf16 transformer stack whatever the trunk's precision: 5 % like the x0-pred rate
    np.testing.assert_allclose(rate.view(-1).cpu().numpy(), z["fwd/rate"], rtol=0.05, atol=1e-4)
    close(am, z["fwd/auto_mean"], 0.03)
    close(asd, z["fwd/auto_std"], 0.03)
    want = ol.trans_forward(packed, z["fwd/x"], z["fwd/onehot"], z["fwd/dims"], z["fwd/ts"], model.forward_rate.as_c(), nearest=z["fwd/nearest"])
    np.testing.assert_allclose(rate.view(-1).cpu().numpy(), want.rate, rtol=0.05, atol=1e-4)
    if precision == "bf16":   # the whole sampler on such a model: births happen, multiplicities stay in range
        sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
        sk["dt"] = 0.05
        sampler = JumpSampler(structure=model.structure, **sk)
        sampler.seed = 3
        in_st = model.make_batch(torch.zeros(64, N, 3, device=DEV), torch.zeros(64, N, S, device=DEV), torch.full((64,), N, device=DEV))
        d = sampler.sample(model.net, in_st, model.jump_diffusion_loss, jet_offset=0).get_dims()
        assert int(d.min()) >= 1 and int(d.max()) > 1 and int(d.max()) <= N


def test_forward_samples_nearest_by_inverse_cdf(fixture):
    z, cfg, model, packed = fixture
    net = model.net
    net.model.precision = "bf16"
    st = model.make_batch(to_dev(z["fwd/x"]), to_dev(z["fwd/onehot"]), to_dev(z["fwd/dims"]))
    out = net(st, to_dev(z["fwd/ts"]), forward_rate=model.forward_rate, predict="eps", nearest_atom=None, sample_nearest_atom=True,
              u_nearest=to_dev(z["fwd/u_near"]))
    near = net.model.last_nearest_atom.cpu().numpy()
    assert (near == z["fwd/nearest_sampled"]).mean() >= 5 / 6      # a draw within the bf16 tolerance of a CDF edge may move
    same = near == z["fwd/nearest_sampled"]
    close(out[2][0].cpu().numpy()[same], z["fwd/auto_mean_sampled"][same], 0.03)


def test_tokens_in_the_underflow_regime(fixture):
    """late sampler states of the fixture hold one-hot values ~1e3: the batch-axis softmax underflows into the denormals and
    a wrong token changes D_xt by O(1).  fp32 trunk vs oracle on those states."""
    z, cfg, model, packed = fixture
    model.net.model.precision = "fp32"
    fr = model.forward_rate.as_c()
    for i in (12, 17, 19):
        x, oh, dims = z["smp/x_traj"][i], z["smp/oh_traj"][i], z["smp/dims_traj"][i]
        B = x.shape[0]
        ts = np.full(B, z["smp/ts"][i], np.float32)
        want = ol.trans_forward(packed, x, oh, dims, ts, fr, u_nearest=z["smp/u_near"][i])
        st = model.make_batch(to_dev(x), to_dev(oh), to_dev(dims))
        D = model.net(st, to_dev(ts), forward_rate=model.forward_rate, nearest_atom=None, sample_nearest_atom=True,
                      u_nearest=to_dev(z["smp/u_near"][i]))[0]
        close(D, want.d_xt, 1e-4)


def test_sampler_update_kernel_matches_oracle():
    """the fused HBM-bound update alone, same inputs on both sides: dims bit-exact, state to fp32 summation-order noise"""
    g = np.random.default_rng(5)
    for (B, N, S) in ((7, 16, 8), (64, 128, 8), (5, 30, 4)):
        F = 3 + S
        dims = g.integers(1, N + 1, B).astype(np.int32)
        dims[0], dims[1] = 1, N
        m = (np.arange(N)[None] < dims[:, None])[..., None]
        x = (g.standard_normal((B, N, 3)) * m).astype(np.float32)
        oh = (g.standard_normal((B, N, S)) * m).astype(np.float32)
        v, lg = g.standard_normal((B, N, 3)).astype(np.float32), g.standard_normal((B, N, S)).astype(np.float32)
        rate = (g.random(B) * 20).astype(np.float32)
        nm, ns = g.standard_normal((B, F)).astype(np.float32) * 3, g.standard_normal((B, F)).astype(np.float32) * 12
        z_diff, u_jump, z_new = (g.standard_normal((B, N * F)).astype(np.float32), g.random(B).astype(np.float32),
                                 g.standard_normal((B, F)).astype(np.float32))
        x[2, 0, 0] = np.nan
        for c_noise in (0.3, 0.0):
            sc = (1.004, 0.009, c_noise, 1.7, 0.05)
            wx, wo, wd = ol.trans_sampler_update(x, oh, dims, v, lg, rate, nm, ns, *sc, z_diff, u_jump, z_new)
            tx, to, td = to_dev(x), to_dev(oh), to_dev(dims)
            _native.trans_sampler_update(tx, to, td, to_dev(v), to_dev(lg), to_dev(rate), to_dev(nm), to_dev(ns), *sc,
                                         z_diff=to_dev(z_diff), u_jump=to_dev(u_jump), z_new=to_dev(z_new))
            assert np.array_equal(td.cpu().numpy(), wd) and (wd > dims).any()
            np.testing.assert_allclose(tx.cpu().numpy(), wx, rtol=1e-5, atol=2e-6)
            np.testing.assert_allclose(to.cpu().numpy(), wo, rtol=1e-5, atol=2e-6)


def test_corrector_update_kernels_match_oracle():
    """Langevin corrector (+ jump corrector) alone, same inputs on both sides: the batch-norm step size, the stale predictor
    mask, births and deaths: dims bit-exact, state to fp32 summation-order noise (sampler.py:258-312)"""
    g = np.random.default_rng(11)
    for (B, N, S) in ((7, 16, 8), (300, 128, 8), (5, 30, 4)):
        F = 3 + S
        dims = g.integers(1, N + 1, B).astype(np.int32)
        dims[0], dims[1] = 1, N
        mask_dims = np.clip(dims + g.integers(-2, 2, B), 1, N).astype(np.int32)   # born / died since the predictor step
        m = (np.arange(N)[None] < dims[:, None])[..., None]
        x = (g.standard_normal((B, N, 3)) * m).astype(np.float32)
        oh = (g.standard_normal((B, N, S)) * m).astype(np.float32)
        v, lg = (g.standard_normal((B, N, 3)) * m).astype(np.float32), (g.standard_normal((B, N, S)) * m).astype(np.float32)
        rate = (g.random(B) * 20).astype(np.float32)
        nm, ns = g.standard_normal((B, F)).astype(np.float32) * 3, g.standard_normal((B, F)).astype(np.float32) * 12
        z_diff, z_new = g.standard_normal((B, N * F)).astype(np.float32), g.standard_normal((B, F)).astype(np.float32)
        u_jump, u_death = g.random(B).astype(np.float32), g.random(B).astype(np.float32)
        for noise_on, jump in ((1, 1), (0, 1), (1, 0)):
            sc = dict(alpha=0.97, noise_on=noise_on, inv_std=1.7, snr=0.2, jump_dt=0.05, jump_corrector=jump, death_prob=0.4)
            wx, wo, wd = ol.trans_corrector_update(x, oh, dims, mask_dims, v, lg, rate, nm, ns, sc["alpha"], noise_on, sc["inv_std"],
                                                   sc["snr"], sc["jump_dt"], jump, sc["death_prob"], z_diff, u_jump, u_death, z_new)
            tx, to, td = to_dev(x), to_dev(oh), to_dev(dims)
            step = _native.trans_corrector_update(tx, to, td, to_dev(v), to_dev(lg), to_dev(rate), to_dev(nm), to_dev(ns), sc["alpha"],
                                                  noise_on, sc["inv_std"], sc["snr"], sc["jump_dt"], bool(jump), sc["death_prob"],
                                                  mask_dims=to_dev(mask_dims), z_diff=to_dev(z_diff), u_jump=to_dev(u_jump),
                                                  u_death=to_dev(u_death), z_new=to_dev(z_new))
            assert 0 < float(step) < 1
            assert np.array_equal(td.cpu().numpy(), wd)
            if jump:
                assert not np.array_equal(wd, dims) and (B < 100 or ((wd > dims).any() and (wd < dims).any()))
            else:
                assert np.array_equal(wd, dims)
            np.testing.assert_allclose(tx.cpu().numpy(), wx, rtol=1e-5, atol=3e-6)
            np.testing.assert_allclose(to.cpu().numpy(), wo, rtol=1e-5, atol=3e-6)
            dead = np.arange(N)[None] >= wd[:, None]
            assert (tx.cpu().numpy()[dead] == 0).all() and (to.cpu().numpy()[dead] == 0).all()


def test_corrector_in_kernel_noise_matches_the_norm_pass():
    """Philox path of a corrector row: the norm kernel and the update kernels regenerate the same draws — with score = 0 and
    snr chosen so that sqrt(2 step) = 1 the state moves by exactly the centred noise whose norm set the step."""
    B, N, S = 256, 128, 8
    dims = torch.full((B,), N, dtype=torch.int32, device=DEV)
    x, oh = torch.zeros(B, N, 3, device=DEV), torch.zeros(B, N, S, device=DEV)
    v, lg = torch.ones(B, N, 3, device=DEV), torch.ones(B, N, S, device=DEV)   # |score_b| = sqrt(N F) for every jet
    step = _native.trans_corrector_update(x, oh, dims, v, lg, None, None, None, 1.0, True, 1.0, 1.0, 0.02, seed=3, jet_offset=40, step=9)
    step = float(step)
    sq, F = (2 * step) ** 0.5, 3 + S
    noise_oh = (oh + step) / sq           # increment = -step * 1 + sqrt(2 step) * z
    assert abs(float(noise_oh.mean())) < 5e-3 and abs(float(noise_oh.std()) - 1.0) < 5e-3
    # step = (noise_norm / grad_norm)^2 * 2 with grad_norm = sqrt(N F): recover noise_norm and compare with the noise actually applied
    noise_norm = (step / 2) ** 0.5 * (N * F) ** 0.5
    # the continuous part lost the (constant) -step score under the centre-of-mass removal: it is the centred noise itself
    applied = torch.cat([(x / sq).flatten(1), noise_oh.flatten(1)], 1)
    assert abs(float(applied.norm(dim=1).mean()) - noise_norm) < 2e-3 * noise_norm


def test_sampler_update_in_kernel_normals_are_standard():
    """Philox + Box-Muller draws of the update kernel: with x = v = 0 and c_noise = 1 the output IS the centred noise."""
    B, N, S = 512, 128, 8
    dims = torch.full((B,), N, dtype=torch.int32, device=DEV)
    x, oh = torch.zeros(B, N, 3, device=DEV), torch.zeros(B, N, S, device=DEV)
    zero = lambda *s: torch.zeros(*s, device=DEV)
    _native.trans_sampler_update(x, oh, dims, zero(B, N, 3), zero(B, N, S), zero(B), zero(B, 3 + S), zero(B, 3 + S), 0.0, 0.0, 1.0, 1.0,
                                 0.02, seed=5, jet_offset=17, step=3)
    z = oh.flatten()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05 and float(z.abs().max()) > 4.0
    assert x.sum(1).abs().max() < 1e-4 and abs(float(x.std()) - (1 - 1 / N) ** 0.5) < 5e-3     # centred over the live particles
    x2, oh2 = torch.zeros_like(x), torch.zeros_like(oh)
    _native.trans_sampler_update(x2[:100], oh2[:100], dims[:100].clone(), zero(100, N, 3), zero(100, N, S), zero(100), zero(100, 3 + S),
                                 zero(100, 3 + S), 0.0, 0.0, 1.0, 1.0, 0.02, seed=5, jet_offset=17 + 50, step=3)
    assert torch.equal(oh2[:50], oh[50:100])       # keyed by the global jet index: shard-invariant draws


def test_tokens_with_many_jets_follow_the_oracle():
    """the chunked column statistics (B > 128) give the oracle's tokens: fp32 trunk output vs oracle on 300 jets"""
    cfg = TransdimensionalEpicConfig()
    cfg.data.max_num_particles = 16
    torch.manual_seed(2)
    model = TransdimensionalJumpDiffusion(cfg).to(DEV)
    model.net.model.precision = "fp32"
    packed = ol.trans_packed(model.cpu())
    model.to(DEV)
    g = np.random.default_rng(3)
    B, N, S = 300, 16, 8
    dims = g.integers(1, N + 1, B).astype(np.int32)
    m = (np.arange(N)[None] < dims[:, None])[..., None]
    x = (g.standard_normal((B, N, 3)) * m).astype(np.float32)
    oh = (g.standard_normal((B, N, S)) * 40 * m).astype(np.float32)     # wide spread: deep in the underflow regime
    ts = g.random(B).astype(np.float32) * 0.99 + 0.005
    near = (g.random(B) * dims).astype(np.int32)
    want = ol.trans_forward(packed, x, oh, dims, ts, model.forward_rate.as_c(), nearest=near)
    D = model.net(model.make_batch(to_dev(x), to_dev(oh), to_dev(dims)), to_dev(ts), forward_rate=model.forward_rate,
                  nearest_atom=to_dev(near))[0]
    close(D, want.d_xt, 1e-4)


def test_sampler_steps_against_reference_trajectory(fixture):
    """every recorded step of the reference run: one native evaluation + update from the reference's state"""
    z, cfg, model, packed = fixture
    m = model.net.model
    dev = torch.device(DEV)
    sched = jump_schedule(float(z["smp/dt"]), model.noise_schedule)
    fr = model.forward_rate.as_c()
    checked = 0
    for i in range(sched.n_steps):
        x, oh, dims = to_dev(z["smp/x_traj"][i]), to_dev(z["smp/oh_traj"][i]), to_dev(z["smp/dims_traj"][i])
        one = SimpleNamespace(n_steps=1, ts=sched.ts[i:i + 1], c_decay=sched.c_decay[i:i + 1], c_score=sched.c_score[i:i + 1],
                              c_noise=sched.c_noise[i:i + 1], inv_std=sched.inv_std[i:i + 1], jump_dt=sched.jump_dt)
        noise = SimpleNamespace(z_diff=to_dev(z["smp/z_diff"][i:i + 1]), u_near=to_dev(z["smp/u_near"][i:i + 1]),
                                u_jump=to_dev(z["smp/u_jump"][i:i + 1]), z_new=to_dev(z["smp/z_new"][i:i + 1]))
        _native.trans_sample(m.native_trunk(dev), m.native_heads(dev), x, oh, dims, one, fr, noise=noise, precision="bf16")
        last = i + 1 == sched.n_steps
        rx, ro, rd = ((z["smp/x_final"], z["smp/oh_final"], z["smp/dims_final"]) if last else
                      (z["smp/x_traj"][i + 1], z["smp/oh_traj"][i + 1], z["smp/dims_traj"][i + 1]))
        # a birth decision is u < rate*dt: skip jets whose draw lies within the rate tolerance of the threshold
        thr = z["smp/rate_traj"][i] * sched.jump_dt
        safe = np.abs(z["smp/u_jump"][i] - thr) > 0.06 * thr
        assert np.array_equal(dims.cpu().numpy()[safe], rd[safe]), f"step {i}"
        # jets whose nearest particle (hence the new particle) could differ are those that gave birth; the others must
        # follow the reference closely
        quiet = safe & (rd == z["smp/dims_traj"][i])
        if not quiet.any():
            continue
        scale = max(1.0, np.abs(rx).max(), np.abs(ro).max())
        assert np.abs(x.cpu().numpy()[quiet] - rx[quiet]).max() <= 0.02 * scale, f"step {i}"
        assert np.abs(oh.cpu().numpy()[quiet] - ro[quiet]).max() <= 0.02 * scale, f"step {i}"
        checked += int(quiet.sum())
    assert checked >= sched.n_steps * 2


def test_jump_sampler_end_to_end_properties():
    """JumpSampler.sample through the public API with in-kernel Philox draws at the transepic shape (N = 128):
    deterministic, multiplicities grow from 1, state finite, dead slots zero, continuous features centred."""
    cfg = TransdimensionalEpicConfig()
    cfg.sampler_kwargs.dt = 0.02
    torch.manual_seed(0)
    model = TransdimensionalJumpDiffusion(cfg).to(DEV)
    N, S = cfg.data.max_num_particles, cfg.data.vocab_size_features
    B = 300
    sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
    sampler = JumpSampler(structure=model.structure, **sk)
    sampler.seed = 9
    in_st = model.make_batch(torch.zeros(B, N, 3, device=DEV), torch.zeros(B, N, S, device=DEV), torch.full((B,), N, device=DEV))
    a = sampler.sample(model.net, in_st, model.jump_diffusion_loss, jet_offset=0)
    b = sampler.sample(model.net, in_st, model.jump_diffusion_loss, jet_offset=0)
    assert torch.equal(a.tuple_batch[0], b.tuple_batch[0]) and torch.equal(a.get_dims(), b.get_dims())
    dims = a.get_dims()
    x, oh = a.tuple_batch
    assert x.is_cuda and dims.dtype == torch.int64 and dims.min() >= 1 and dims.max() <= N and dims.float().mean() > 1.5
    assert torch.isfinite(x).all() and torch.isfinite(oh).all()
    dead = torch.arange(N, device=DEV)[None] >= dims[:, None]
    assert (x[dead] == 0).all() and (oh[dead] == 0).all()
    com = x.sum(1) / dims[:, None]
    assert com.abs().max() <= 1e-3 * max(1.0, float(x.abs().max()))


def test_jump_sampler_c_time_grid_runs_the_reference_grid(fixture):
    """dt_schedule='C' + no_noise_final_step through the public API, injected draws: same number of evaluations as the
    reference run, and the jets that never came near a birth threshold end where the reference's did."""
    z, cfg, model, packed = fixture
    B, N, S = z["smpC/oh_final"].shape
    sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
    sk.update(dt=float(z["smp/dt"]), dt_schedule="C", dt_schedule_h=0.1, dt_schedule_l=0.04, dt_schedule_tc=0.5, no_noise_final_step=True)
    sampler = JumpSampler(structure=model.structure, **sk)
    n = len(z["smpC/ts"])
    noise = SimpleNamespace(z_init=to_dev(z["smp/z_init"]), z_diff=to_dev(z["smp/z_diff"][:n]), u_near=to_dev(z["smp/u_near"][:n]),
                            u_jump=to_dev(z["smp/u_jump"][:n]), z_new=to_dev(z["smp/z_new"][:n]))
    in_st = model.make_batch(torch.zeros(B, N, 3, device=DEV), torch.zeros(B, N, S, device=DEV), torch.full((B,), N, device=DEV))
    out = sampler.sample(model.net, in_st, model.jump_diffusion_loss, noise=noise, precision="fp32")
    assert sampler.last_n_steps == n
    dims = out.get_dims().cpu().numpy()
    assert (np.abs(dims - z["smpC/dims_final"]) <= 1).all() and (dims == z["smpC/dims_final"]).sum() >= B - 1
    same = dims == z["smpC/dims_final"]
    x = out.tuple_batch[0].cpu().numpy()
    assert np.isfinite(x).all()
    scale = max(1.0, np.abs(z["smpC/x_final"]).max())
    assert np.median(np.abs(x[same] - z["smpC/x_final"][same])) <= 0.02 * scale


def test_jump_sampler_with_correctors_follows_the_reference_groups(fixture):
    """corrector_steps = 2 + jump corrector + no_noise_final_step: every predictor step of the reference run together with
    its corrector rows, restarted from the reference's recorded state (fp32 trunk), ends where the reference's did"""
    import json
    z, cfg, model, packed = fixture
    m = model.net.model
    dev = torch.device(DEV)
    kw = json.loads(str(z["smpL/kwargs"]))
    sched = jump_schedule(float(z["smpL/dt"]), model.noise_schedule, kw["no_noise_final_step"], corrector_steps=kw["corrector_steps"],
                          corrector_snr=kw["corrector_snr"], corrector_start_time=kw["corrector_start_time"],
                          corrector_finish_time=kw["corrector_finish_time"], do_jump_corrector=kw["do_jump_corrector"],
                          forward_rate=model.forward_rate)
    assert sched.n_steps == len(z["smpL/ts"])
    starts = [i for i in range(sched.n_steps) if sched.kind[i] == 0] + [sched.n_steps]
    fr = model.forward_rate.as_c()
    same = total = 0
    errs = []
    for a, b in zip(starts[:-1], starts[1:]):
        if b - a == 1:
            continue   # predictor-only rows are covered by test_sampler_steps_against_reference_trajectory
        x, oh, dims = to_dev(z["smpL/x_traj"][a]), to_dev(z["smpL/oh_traj"][a]), to_dev(z["smpL/dims_traj"][a])
        sub = SimpleNamespace(n_steps=b - a, kind=sched.kind[a:b], ts=sched.ts[a:b], c_decay=sched.c_decay[a:b], c_score=sched.c_score[a:b],
                              c_noise=sched.c_noise[a:b], inv_std=sched.inv_std[a:b], death_prob=sched.death_prob[a:b],
                              jump_dt=sched.jump_dt, corrector_snr=sched.corrector_snr, jump_corrector=sched.jump_corrector)
        noise = SimpleNamespace(z_diff=to_dev(z["smpL/z_diff"][a:b]), u_near=to_dev(z["smpL/u_near"][a:b]),
                                u_jump=to_dev(z["smpL/u_jump"][a:b]), z_new=to_dev(z["smpL/z_new"][a:b]),
                                u_death=to_dev(z["smpL/u_death"][a:b]))
        _native.trans_sample(m.native_trunk(dev), m.native_heads(dev), x, oh, dims, sub, fr, noise=noise, precision="fp32")
        last = b == sched.n_steps
        rx, rd = ((z["smpL/x_final"], z["smpL/dims_final"]) if last else (z["smpL/x_traj"][b], z["smpL/dims_traj"][b]))
        got = dims.cpu().numpy()
        total += len(rd)
        same += int((got == rd).sum())
        quiet = (got == rd) & (rd == z["smpL/dims_traj"][a])
        if quiet.any():
            errs.append(np.abs(x.cpu().numpy()[quiet] - rx[quiet]).max() / max(1.0, np.abs(rx).max()))
    assert total >= 50 and same >= 0.9 * total, (same, total)
    assert len(errs) >= 5 and np.median(errs) <= 0.02, errs


def test_jump_sampler_with_correctors_philox_properties():
    """the same configuration through JumpSampler.sample with in-kernel draws at N = 128: deterministic, finite,
    dead slots zero, centred (the Langevin step size is a statistic of the whole batch, as in the reference)"""
    cfg = TransdimensionalEpicConfig()
    torch.manual_seed(0)
    model = TransdimensionalJumpDiffusion(cfg).to(DEV)
    N, S, B = cfg.data.max_num_particles, cfg.data.vocab_size_features, 200
    sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
    sk.update(dt=0.02, corrector_steps=1, corrector_snr=0.1, corrector_start_time=0.5, corrector_finish_time=0.05, do_jump_corrector=True)
    sampler = JumpSampler(structure=model.structure, **sk)
    sampler.seed = 4
    start = lambda n: model.make_batch(torch.zeros(n, N, 3, device=DEV), torch.zeros(n, N, S, device=DEV), torch.full((n,), N, device=DEV))
    a = sampler.sample(model.net, start(B), model.jump_diffusion_loss, jet_offset=0)
    assert sampler.last_n_steps > 50
    b = sampler.sample(model.net, start(B), model.jump_diffusion_loss, jet_offset=0)
    assert torch.equal(a.tuple_batch[0], b.tuple_batch[0]) and torch.equal(a.get_dims(), b.get_dims())
    dims, (x, oh) = a.get_dims(), a.tuple_batch
    assert dims.min() >= 1 and dims.max() <= N and torch.isfinite(x).all() and torch.isfinite(oh).all()
    dead = torch.arange(N, device=DEV)[None] >= dims[:, None]
    assert (x[dead] == 0).all() and (oh[dead] == 0).all()
    assert (x.sum(1) / dims[:, None]).abs().max() <= 1e-3 * max(1.0, float(x.abs().max()))


def test_c3_full_size_properties():
    """BASELINE config 3 shape (B = 8192, N = 128) — size-independent properties of one evaluation and of a sampler run:
    the network treats a jet the same wherever it sits in the batch (tokens use per-column batch statistics, which a
    duplicated jet shares), nothing is written outside the live slot of a birth, a full jet cannot give birth, and the
    sampler keeps its invariants (dead slots zero, centred continuous features, multiplicities in range and growing)."""
    cfg = TransdimensionalEpicConfig()
    torch.manual_seed(4)
    model = TransdimensionalJumpDiffusion(cfg).to(DEV)
    B, N, S = 8192, cfg.data.max_num_particles, cfg.data.vocab_size_features
    g = torch.Generator().manual_seed(5)
    dims = torch.randint(1, N + 1, (B,), generator=g)
    dims[:64] = N
    m = (torch.arange(N)[None] < dims[:, None]).float().unsqueeze(-1)
    x, oh = torch.randn(B, N, 3, generator=g) * m, torch.randn(B, N, S, generator=g) * m
    ts = torch.rand(B, generator=g) * 0.99 + 0.005
    near = (torch.rand(B, generator=g) * dims).long()
    src = torch.arange(100, 164)
    dst = torch.arange(5000, 5064)     # copies of 64 jets far away in the batch (a different CTA, a different position in its queue)
    for t in (x, oh, dims, ts, near):
        t[dst] = t[src]
    out = model.net(model.make_batch(x.to(DEV), oh.to(DEV), dims.to(DEV)), ts.to(DEV), forward_rate=model.forward_rate,
                    nearest_atom=near.to(DEV))
    D, rate, (am, asd), x0l, nal = [o.cpu() if torch.is_tensor(o) else tuple(t.cpu() for t in o) for o in out]
    for t in (D, rate, am, asd, x0l, nal):
        assert torch.isfinite(t).all() and torch.equal(t[dst], t[src])
    assert (rate[:64] == 0).all() and (am[:64] == 0).all()                    # full jets: no slot to fill
    F = 3 + S
    slot = torch.zeros(B, N * F, dtype=torch.bool)
    rows = torch.arange(B)[dims < N]
    for c in range(3):
        slot[rows, dims[rows] * 3 + c] = True
    for s_ in range(S):
        slot[rows, N * 3 + dims[rows] * S + s_] = True
    assert (am[~slot] == 0).all() and (asd[~slot] == 0).all() and (rate >= 0).all()
    cfg.sampler_kwargs.dt = 0.1
    sk = {k: v for k, v in vars(cfg.sampler_kwargs).items() if k not in ("class_name", "do_jump_back", "jump_back_start_time")}
    sampler = JumpSampler(structure=model.structure, **sk)
    in_st = model.make_batch(torch.zeros(B, N, 3, device=DEV), torch.zeros(B, N, S, device=DEV), torch.full((B,), N, device=DEV))
    st = sampler.sample(model.net, in_st, model.jump_diffusion_loss, jet_offset=0)
    d, (xs, ohs) = st.get_dims(), st.tuple_batch
    assert d.min() >= 1 and d.max() <= N and d.float().mean() > 1.2 and torch.isfinite(xs).all() and torch.isfinite(ohs).all()
    dead = torch.arange(N, device=DEV)[None] >= d[:, None]
    assert (xs[dead] == 0).all() and (ohs[dead] == 0).all()
    assert (xs.sum(1) / d[:, None]).abs().max() <= 1e-3 * max(1.0, float(xs.abs().max()))


def test_trans_errors_are_loud(fixture):
    z, cfg, model, packed = fixture
    m = model.net.model
    dev = torch.device(DEV)
    st = model.make_batch(torch.from_numpy(z["fwd/x"]), torch.from_numpy(z["fwd/onehot"]), torch.from_numpy(z["fwd/dims"]))
    with pytest.raises(_native.MmbError):
        model.net(st, torch.from_numpy(z["fwd/ts"]), forward_rate=model.forward_rate, nearest_atom=torch.zeros(6).long())   # host tensors
    with pytest.raises(_native.MmbError):
        _native.TransHeads(m.trans_dims(), torch.zeros(10), dev)
