"""Golden fixture of the trans-dimensional jump diffusion, produced by the reference (container only):

    python tests/golden/make_golden_trans.py

* ``forward``: EpsilonPrecond.forward -> TransdimensionalEPiC.forward
  (mp/models/generative/transdimensional/transdimensional_model.py:124-133, 245-426) on a small batch with
  the nearest particle (a) given and (b) sampled through a caller-supplied ``rnd`` (inverse CDF on an injected
  uniform instead of torch.multinomial's generator stream).
* ``sample``: JumpSampler.sample (sampler.py:157-324) for 20 steps with every draw injected through ``rnd``.
  The function does not run as shipped: the two documented one-line patches are applied here —
  ``gs.max_problem_dim = max_num_particles`` and EpsilonPrecond.forward passing ``sample_nearest_atom``/``rnd``
  on to the model.  Nothing else of the reference is modified.

The data module is built from an in-memory fake dataset (real HDF5 files cannot be read here).
"""
import contextlib
import io
import json
import os
import sys
import warnings
from dataclasses import asdict
from types import SimpleNamespace

import numpy as np
import torch

warnings.simplefilter("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shim)

from multimodal_particles.config_classes.transdimensional_unconditional_config import TransdimensionalEpicConfig  # noqa: E402
from multimodal_particles.models.generative.transdimensional.transdimensional_model import (  # noqa: E402
    EpsilonPrecond, TransdimensionalJumpDiffusion)
from multimodal_particles.models.generative.transdimensional.structure import StructuredDataBatch  # noqa: E402
from multimodal_particles.models.generative.transdimensional.sampler import JumpSampler  # noqa: E402
from multimodal_particles.data.particle_clouds.jets_dataloader import JetsDataloaderModule  # noqa: E402

quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # structure.py:156 prints a shape per get_mask call


def fake_datamodule(cfg, n_jets=40, seed=0):
    N, S = cfg.data.max_num_particles, cfg.data.vocab_size_features
    g = torch.Generator().manual_seed(seed)
    mult = torch.randint(1, N + 1, (n_jets,), generator=g)
    mask = (torch.arange(N)[None] < mult[:, None]).long().unsqueeze(-1)

    class Target(SimpleNamespace):
        def __len__(self):
            return n_jets

    target = Target(continuous=torch.randn(n_jets, N, 3, generator=g) * mask,
                    discrete=torch.randint(0, S, (n_jets, N, 1), generator=g) * mask, mask=mask)
    fake = SimpleNamespace(config=cfg, vocab_size_features=S, vocab_size_context=0, target=target, source=SimpleNamespace())
    with quiet():
        return JetsDataloaderModule(cfg, jetdataset=fake)


class InjectedRnd:
    """Caller-supplied ``rnd`` of JumpSampler.sample / TransdimensionalEPiC.forward with pre-drawn values.
    Call order inside one sampler step: multinomial (network), randn_like(xt), rand, randn_like(std)."""

    def __init__(self, N, S, z_init=None, z_diff=None, u_near=None, u_jump=None, z_new=None):
        self.N, self.S = N, S
        self.z_init, self.z_diff, self.u_near, self.u_jump, self.z_new = z_init, z_diff, u_near, u_jump, z_new
        self.n_randn = self.n_rand = self.n_multi = 0
        self.nearest = []

    def randn_like(self, t):
        i, self.n_randn = self.n_randn, self.n_randn + 1
        if self.z_init is not None:
            if i == 0:
                return self.z_init.clone()
            i -= 1
        step, which = divmod(i, 2)
        if which == 0:
            return self.z_diff[step].clone()
        z = self.z_new[step]  # [B, 3+S] -> the same draw offered to every slot; only slot `dims` is kept
        B = z.shape[0]
        return torch.cat([z[:, None, :3].expand(B, self.N, 3).reshape(B, -1),
                          z[:, None, 3:].expand(B, self.N, self.S).reshape(B, -1)], 1).clone()

    def rand(self, size, device=None):
        i, self.n_rand = self.n_rand, self.n_rand + 1
        return self.u_jump[i].clone()

    def multinomial(self, probs, num_samples=1):
        i, self.n_multi = self.n_multi, self.n_multi + 1
        c = torch.cumsum(probs, 1)
        idx = (self.u_near[i][:, None] >= c).sum(1).clamp(max=probs.shape[1] - 1)
        self.nearest.append(idx.clone())
        return idx.view(-1, 1)


class EvalRnd:
    """``rnd`` indexed by NETWORK EVALUATION (every evaluation starts with the nearest-particle multinomial): within
    evaluation e the first randn_like is the diffusion/Langevin noise z_diff[e], the second the new particle z_new[e];
    the first rand is the birth uniform u_jump[e], the second the death uniform u_death[e] (jump corrector only)."""

    def __init__(self, N, S, z_init, z_diff, u_near, u_jump, u_death, z_new):
        self.N, self.S = N, S
        self.z_init, self.z_diff, self.u_near, self.u_jump, self.u_death, self.z_new = z_init, z_diff, u_near, u_jump, u_death, z_new
        self.e, self.first, self.n_randn, self.n_rand = -1, True, 0, 0

    def multinomial(self, probs, num_samples=1):
        self.e, self.n_randn, self.n_rand = self.e + 1, 0, 0
        c = torch.cumsum(probs, 1)
        return (self.u_near[self.e][:, None] >= c).sum(1).clamp(max=probs.shape[1] - 1).view(-1, 1)

    def randn_like(self, t):
        if self.first:
            self.first = False
            return self.z_init.clone()
        i, self.n_randn = self.n_randn, self.n_randn + 1
        if i == 0:
            return self.z_diff[self.e].clone()
        z = self.z_new[self.e]
        B = z.shape[0]
        return torch.cat([z[:, None, :3].expand(B, self.N, 3).reshape(B, -1),
                          z[:, None, 3:].expand(B, self.N, self.S).reshape(B, -1)], 1).clone()

    def rand(self, size, device=None):
        i, self.n_rand = self.n_rand, self.n_rand + 1
        return (self.u_jump if i == 0 else self.u_death)[self.e].clone()


def patched_precond_forward(self, st_batch, ts, predict='eps', forward_rate=None, nearest_atom=None, **kw):
    """EpsilonPrecond.forward (transdimensional_model.py:124-133) with the missing kwargs passed through."""
    eps, *others = self.model(st_batch, ts, nearest_atom=nearest_atom, forward_rate=forward_rate, **kw)
    assert predict == 'eps'
    return eps, *others


def main():
    cfg = TransdimensionalEpicConfig()
    cfg.data.return_type = "list"
    N = cfg.data.max_num_particles = 16
    S, F = cfg.data.vocab_size_features, 3 + cfg.data.vocab_size_features
    dm = fake_datamodule(cfg)
    torch.manual_seed(301)
    with quiet():
        model = TransdimensionalJumpDiffusion(cfg, dm)
    net = model.net
    with torch.no_grad():  # make rates, nearest-particle choices and birth statistics non-trivial at random init
        net.model.post_rate_proj.weight.mul_(4.0)
        net.model.near_atom_proj.weight.mul_(8.0)
        net.model.vec_weighting_proj.weight.mul_(4.0)
        net.model.post_auto_proj.weight.mul_(4.0)
        net.model.epic.epic.output_layer.weight_g.mul_(3.0)
    gs = dm.graphical_structure
    gs.max_problem_dim = N   # patch (i) of SURVEY §3.3
    EpsilonPrecond.forward = patched_precond_forward   # patch (ii)

    out = dict(config=json.dumps(asdict(cfg)))
    g = torch.Generator().manual_seed(302)

    # ---- forward fixture
    B = 6
    dims = torch.tensor([1, 3, 16, 8, 2, 5])
    m = (torch.arange(N)[None] < dims[:, None]).float().unsqueeze(-1)
    x = torch.randn(B, N, 3, generator=g) * m
    x = x - (x.sum(1, keepdim=True) / dims.view(B, 1, 1)) * m
    oh = torch.randn(B, N, S, generator=g) * m
    ts = torch.tensor([0.05, 0.5, 0.999, 0.2, 0.75, 0.011])
    nearest = torch.tensor([0, 2, 7, 0, 1, 4])
    u_near = torch.rand(1, B, generator=g)
    st = lambda: StructuredDataBatch([x.clone(), oh.clone()], dims.clone(), dm.observed, dm.exist, dm.is_onehot, gs)
    with quiet(), torch.no_grad():
        D, rate, (am, asd), x0l, nal = net(st(), ts, forward_rate=model.forward_rate, predict="eps", nearest_atom=nearest)
        rnd = InjectedRnd(N, S, u_near=u_near)
        D2, rate2, (am2, asd2), x0l2, nal2 = net(st(), ts, forward_rate=model.forward_rate, predict="eps", nearest_atom=None,
                                                 sample_nearest_atom=True, rnd=rnd)
        tokens = st().from_st_batch_to_multimodal_bridge_databatch()[1]
    assert torch.equal(D, D2) and torch.equal(nal, nal2)
    out.update({"fwd/x": x.numpy(), "fwd/onehot": oh.numpy(), "fwd/dims": dims.numpy().astype(np.int32), "fwd/ts": ts.numpy(),
                "fwd/nearest": nearest.numpy().astype(np.int32), "fwd/u_near": u_near[0].numpy(),
                "fwd/tokens": tokens[..., 0].numpy().astype(np.uint8),
                "fwd/d_xt": D.numpy(), "fwd/rate": rate.view(-1).numpy(), "fwd/auto_mean": am.numpy(), "fwd/auto_std": asd.numpy(),
                "fwd/x0_dim_logits": x0l.numpy(), "fwd/near_atom_logits": nal.numpy(),
                "fwd/nearest_sampled": rnd.nearest[0].numpy().astype(np.int32),
                "fwd/auto_mean_sampled": am2.numpy(), "fwd/auto_std_sampled": asd2.numpy(), "fwd/rate_sampled": rate2.view(-1).numpy()})
    fr = model.forward_rate
    out["forward_rate"] = np.array([0.0, fr.get_scalar(), fr.offset, fr.rate_cut_t], dtype=np.float64)

    # ---- sampler fixture: dt = 0.05 -> 20 steps
    dt, steps, Bs = 0.05, 20, 5
    sk = asdict(cfg.sampler_kwargs)
    sk.update(dt=dt)
    for key in ("class_name", "do_jump_back", "jump_back_start_time"):
        sk.pop(key)
    sampler = JumpSampler(structure=model.structure, **sk)
    z_init = torch.randn(Bs, N * F, generator=g)
    z_diff = torch.randn(steps, Bs, N * F, generator=g)
    u_near_s = torch.rand(steps, Bs, generator=g)
    u_jump = torch.rand(steps, Bs, generator=g)
    z_new = torch.randn(steps, Bs, F, generator=g)
    rnd = InjectedRnd(N, S, z_init=z_init, z_diff=z_diff, u_near=u_near_s, u_jump=u_jump, z_new=z_new)
    rec = dict(ts=[], dims=[], x=[], oh=[], rate=[], d_xt=[])
    orig_get_score = sampler.get_score

    def get_score(state_st_batch, net_, loss, ts_, dataset_obj, rnd_):
        rec["ts"].append(ts_[0].item())
        rec["dims"].append(state_st_batch.get_dims().clone())
        rec["x"].append(state_st_batch.tuple_batch[0].clone())
        rec["oh"].append(state_st_batch.tuple_batch[1].clone())
        score, rate_xt, mean_std = orig_get_score(state_st_batch, net_, loss, ts_, dataset_obj, rnd_)
        rec["rate"].append(rate_xt.view(-1).clone())
        return score, rate_xt, mean_std

    sampler.get_score = get_score
    in_st = StructuredDataBatch([torch.zeros(Bs, N, 3), torch.zeros(Bs, N, S)], torch.full((Bs,), N), dm.observed, dm.exist,
                                dm.is_onehot, gs)
    with quiet(), torch.no_grad():
        final = sampler.sample(net, in_st, model.jump_diffusion_loss, rnd)
    assert len(rec["ts"]) == steps and rnd.n_multi == steps and rnd.n_rand == steps, (len(rec["ts"]), rnd.n_multi, rnd.n_rand)
    out.update({"smp/dt": np.float64(dt), "smp/z_init": z_init.numpy(), "smp/z_diff": z_diff.numpy(), "smp/u_near": u_near_s.numpy(),
                "smp/u_jump": u_jump.numpy(), "smp/z_new": z_new.numpy(), "smp/ts": np.array(rec["ts"], np.float32),
                "smp/dims_traj": torch.stack(rec["dims"]).numpy().astype(np.int32), "smp/x_traj": torch.stack(rec["x"]).numpy(),
                "smp/oh_traj": torch.stack(rec["oh"]).numpy(), "smp/rate_traj": torch.stack(rec["rate"]).numpy(),
                "smp/nearest_traj": torch.stack(rnd.nearest).numpy().astype(np.int32),
                "smp/x_final": final.tuple_batch[0].numpy(), "smp/oh_final": final.tuple_batch[1].numpy(),
                "smp/dims_final": final.get_dims().numpy().astype(np.int32)})
    # ---- same sampler on the 'C' time grid (sampler.py:79-88) with no_noise_final_step: only the grid and the last step change
    skc = dict(sk, dt_schedule="C", dt_schedule_h=0.1, dt_schedule_l=0.04, dt_schedule_tc=0.5, no_noise_final_step=True)
    sampler_c = JumpSampler(structure=model.structure, **skc)
    rnd_c = InjectedRnd(N, S, z_init=z_init, z_diff=z_diff, u_near=u_near_s, u_jump=u_jump, z_new=z_new)
    ts_c = []
    orig_c = sampler_c.get_score
    sampler_c.get_score = lambda st_, net_, loss_, ts_, d_, r_: (ts_c.append(ts_[0].item()), orig_c(st_, net_, loss_, ts_, d_, r_))[1]
    in_st_c = StructuredDataBatch([torch.zeros(Bs, N, 3), torch.zeros(Bs, N, S)], torch.full((Bs,), N), dm.observed, dm.exist,
                                  dm.is_onehot, gs)
    with quiet(), torch.no_grad():
        final_c = sampler_c.sample(net, in_st_c, model.jump_diffusion_loss, rnd_c)
    assert len(ts_c) <= steps
    out.update({"smpC/ts": np.array(ts_c, np.float32), "smpC/x_final": final_c.tuple_batch[0].numpy(),
                "smpC/oh_final": final_c.tuple_batch[1].numpy(), "smpC/dims_final": final_c.get_dims().numpy().astype(np.int32)})
    # ---- Langevin corrector steps + jump corrector (sampler.py:258-312), dt = 0.02, correctors below t = 0.25 down to the end,
    #      no_noise_final_step: 50 predictor rows + 24 corrector rows
    dtl, rows = 0.02, 74
    skl = dict(sk, dt=dtl, corrector_steps=2, corrector_snr=0.2, corrector_start_time=0.25, corrector_finish_time=0.0,
               do_jump_corrector=True, no_noise_final_step=True)
    sampler_l = JumpSampler(structure=model.structure, **skl)
    gl = torch.Generator().manual_seed(303)
    zl_init = torch.randn(Bs, N * F, generator=gl)
    zl_diff = torch.randn(rows, Bs, N * F, generator=gl)
    ul_near, ul_jump, ul_death = (torch.rand(rows, Bs, generator=gl) for _ in range(3))
    zl_new = torch.randn(rows, Bs, F, generator=gl)
    rnd_l = EvalRnd(N, S, zl_init, zl_diff, ul_near, ul_jump, ul_death, zl_new)
    recl = dict(ts=[], dims=[], x=[], oh=[])
    orig_l = sampler_l.get_score

    def get_score_l(state_st_batch, net_, loss, ts_, dataset_obj, rnd_):
        recl["ts"].append(ts_[0].item())
        recl["dims"].append(state_st_batch.get_dims().clone())
        recl["x"].append(state_st_batch.tuple_batch[0].clone())
        recl["oh"].append(state_st_batch.tuple_batch[1].clone())
        return orig_l(state_st_batch, net_, loss, ts_, dataset_obj, rnd_)

    sampler_l.get_score = get_score_l
    in_st_l = StructuredDataBatch([torch.zeros(Bs, N, 3), torch.zeros(Bs, N, S)], torch.full((Bs,), N), dm.observed, dm.exist,
                                  dm.is_onehot, gs)
    with quiet(), torch.no_grad():
        final_l = sampler_l.sample(net, in_st_l, model.jump_diffusion_loss, rnd_l)
    assert len(recl["ts"]) == rows, len(recl["ts"])
    dl = torch.stack(recl["dims"])
    print("corrector run: dims per row", dl[:, 0].tolist(), "final", final_l.get_dims().tolist())
    out.update({"smpL/dt": np.float64(dtl), "smpL/kwargs": json.dumps({k: skl[k] for k in ("corrector_steps", "corrector_snr",
                "corrector_start_time", "corrector_finish_time", "do_jump_corrector", "no_noise_final_step")}),
                "smpL/z_init": zl_init.numpy(), "smpL/z_diff": zl_diff.numpy(), "smpL/u_near": ul_near.numpy(),
                "smpL/u_jump": ul_jump.numpy(), "smpL/u_death": ul_death.numpy(), "smpL/z_new": zl_new.numpy(),
                "smpL/ts": np.array(recl["ts"], np.float32), "smpL/dims_traj": dl.numpy().astype(np.int32),
                "smpL/x_traj": torch.stack(recl["x"]).numpy(), "smpL/oh_traj": torch.stack(recl["oh"]).numpy(),
                "smpL/x_final": final_l.tuple_batch[0].numpy(), "smpL/oh_final": final_l.tuple_batch[1].numpy(),
                "smpL/dims_final": final_l.get_dims().numpy().astype(np.int32)})
    out.update(mg.np_state_dict(model))
    path = os.path.join(HERE, "trans.npz")
    np.savez_compressed(path, **out)
    print(f"trans: {os.path.getsize(path) / 1024:.0f} KiB; forward rates {rate.view(-1).tolist()}; sampled nearest "
          f"{rnd.nearest[0].tolist() if False else out['fwd/nearest_sampled'].tolist()}; final dims {final.get_dims().tolist()}")


if __name__ == "__main__":
    main()
