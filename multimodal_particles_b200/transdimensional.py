"""Trans-dimensional jump diffusion — drop-in for the generation side of
mp/models/generative/transdimensional/{transdimensional_model,structure,sampler}.py and the parts of
mp/models/generative/diffusion/noising.py and mp/data/particle_clouds/jets_dataloader.py:336-478 they use.

Same class names, constructor arguments, call signatures, tensor layouts and state-dict keys
(``net.model.epic.*``, ``net.model.fc_layer.*``, ``net.model.{temb_net,transformer_1_proj_in,attn_blocks,res_blocks,
pre_rate_proj,post_rate_proj,near_atom_proj,vec_transformer_in_proj,vec_attn_blocks,vec_res_blocks,vec_weighting_proj,
pre_auto_proj,post_auto_proj}``).  The modules only hold parameters; evaluation goes through libmmbridge.so:
``mmb_trans_forward`` for one network evaluation, ``mmb_trans_sample`` for the whole JumpSampler loop.

Differences from the reference, all deliberate:
* ``JumpSampler.sample`` does not run as shipped on jets (``gs.max_problem_dim`` missing,
  ``EpsilonPrecond.forward`` drops ``sample_nearest_atom``/``rnd``); here both are in place.
* draws: ``rnd.multinomial`` / ``rnd.rand`` / ``rnd.randn_like`` are replaced by in-kernel Philox4x32-10 keyed by the
  global jet index (results do not depend on how jets are sharded), or by injected arrays for parity runs.
* the sampler covers the live configuration, the 'C' time grid, ``no_noise_final_step``, Langevin corrector steps and the
  jump corrector; conditioning/guidance is not built (it differentiates through the network).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

from . import _native
from .absorbing_flows import AttnBlock, ResnetBlock
from .epic import EPiCWrapper
from .multimodal_bridge_matching import _ModuleBase
from .sharding import next_jet_offset


# ---- noising.py ------------------------------------------------------------------------------------
class VP_SDE:
    """noising.py:7-39 (the statistics the sampler reads)."""

    def __init__(self, max_dim, beta_min, beta_max):
        self.max_dim, self.beta_min, self.beta_max = max_dim, beta_min, beta_max

    def get_beta_t(self, ts):
        return (ts * self.beta_max + (1 - ts) * self.beta_min).view(-1, 1).repeat(1, self.max_dim)

    def get_sigma(self, times):
        log_term = -0.25 * times ** 2 * (self.beta_max - self.beta_min) - 0.5 * times * self.beta_min
        return torch.sqrt(1 - torch.exp(2 * log_term))

    def get_p0t_stats(self, st_batch, times):
        minibatch = st_batch.get_flat_lats()
        log_term = -0.25 * times ** 2 * (self.beta_max - self.beta_min) - 0.5 * times * self.beta_min
        log_term = log_term.view(minibatch.shape[0], *([1] * (len(minibatch.shape) - 1)))
        return torch.exp(log_term) * minibatch, torch.sqrt(1 - torch.exp(2. * log_term)).expand(*minibatch.shape)


class StateIndependentForwardRate:
    """noising.py:87-121."""

    def __init__(self, max_dim):
        self.max_dim = max_dim
        self.max_num_deletions = max_dim - 1
        self.std_mult = 0.7
        self.offset = 0.1

    def as_c(self) -> "_native.ForwardRate":
        raise NotImplementedError


class StepForwardRate(StateIndependentForwardRate):
    """noising.py:123-141."""

    def __init__(self, max_dim, rate_cut_t):
        super().__init__(max_dim)
        assert 0 < rate_cut_t < 1
        self.rate_cut_t = rate_cut_t

    def get_scalar(self):
        T, c = self.rate_cut_t, self.max_num_deletions
        return (2 * (1 - T) * c + self.std_mult ** 2 * (1 - T) + math.sqrt(
            (-2 * (1 - T) * c - self.std_mult ** 2 * (1 - T)) ** 2 - 4 * (1 - T) ** 2 * c ** 2)) / (2 * (1 - T) ** 2)

    def get_rate(self, dims, ts):
        return self.get_scalar() * (ts > self.rate_cut_t).long() + self.offset

    def get_rate_integral(self, ts):
        T = self.rate_cut_t
        return (ts - T) * self.get_scalar() * (ts > T).long() + self.offset * ts

    def as_c(self):
        return _native.ForwardRate(0, self.get_scalar(), self.offset, self.rate_cut_t)


class ConstForwardRate(StateIndependentForwardRate):
    """noising.py:143-164."""

    def __init__(self, max_dim, scalar=None):
        super().__init__(max_dim)
        self.scalar = scalar

    def get_scalar(self):
        if self.scalar is not None:
            return self.scalar
        c = self.max_num_deletions
        return (2 * c + self.std_mult ** 2 + math.sqrt((self.std_mult ** 2 + 2 * c) ** 2 - 4 * c ** 2)) / 2

    def get_rate(self, dims, ts):
        return self.get_scalar() * torch.ones_like(ts)

    def get_rate_integral(self, ts):
        return self.get_scalar() * ts

    def as_c(self):
        return _native.ForwardRate(1, self.get_scalar(), 0.0, 0.0)


def get_forward_rate(rate_function_name, max_problem_dim, rate_cut_t):
    if rate_function_name == 'step':
        return StepForwardRate(max_problem_dim, rate_cut_t)
    if rate_function_name == 'const':
        return ConstForwardRate(max_problem_dim, None)
    raise ValueError(rate_function_name)


def get_noise_schedule(noise_schedule_name, max_problem_dim, vp_sde_beta_min, vp_sde_beta_max):
    if noise_schedule_name == 'vp_sde':
        return VP_SDE(max_problem_dim, vp_sde_beta_min, vp_sde_beta_max)
    raise ValueError(noise_schedule_name)


# ---- structure.py / jets_dataloader.py ------------------------------------------------------------------
class JetsGraphicalStructure:
    """The fields and the two batch operations of jets_dataloader.py:336-478 that the generation path uses.
    Built from the config alone (``return_type='list'`` batches: target_continuous, target_discrete one-hot)."""

    def __init__(self, datamodule_or_config):
        config = getattr(datamodule_or_config, "config", datamodule_or_config)
        d = config.data
        self.names_in_batch = ["target_continuous", "target_discrete"]
        self.name_to_index = {"target_continuous": 0, "target_discrete": 1}
        self.max_num_particles = self.max_problem_dim = d.max_num_particles
        self.num_jets = d.num_jets
        self.dim_features_continuous, self.dim_features_discrete = d.dim_features_continuous, d.dim_features_discrete
        self.vocab_size_features, self.vocab_size_context = d.vocab_size_features, d.vocab_size_context
        self.observed, self.exist, self.is_onehot = np.array([0, 0]), np.array([1, 1]), np.array([0, 1])
        self.with_onehot_shapes = [torch.Size([d.max_num_particles, d.dim_features_continuous]),
                                   torch.Size([d.max_num_particles, d.vocab_size_features])]

    def shapes_with_onehot(self):
        return self.with_onehot_shapes

    def remove_problem_dims(self, data, new_dims):
        out = []
        for t in data:
            keep = torch.arange(t.shape[1], device=t.device).view(1, -1, 1) < new_dims.to(t.device).view(-1, 1, 1)
            out.append(t * keep)
        return out

    def adjust_st_batch(self, st_batch):
        """nan_to_num + centre-of-mass removal over the live particles (jets_dataloader.py:433-478)."""
        x, oh = st_batch.tuple_batch[0], st_batch.tuple_batch[1]
        dims = st_batch.get_dims().to(x.device)
        x, oh = torch.nan_to_num(x), torch.nan_to_num(oh)
        node = (torch.arange(x.shape[1], device=x.device).view(1, -1) < dims.view(-1, 1)).long().unsqueeze(2)
        node[dims == 0] = 1
        mean = torch.sum(x, dim=1, keepdim=True) / node.sum(1, keepdim=True)
        st_batch.tuple_batch = (x - mean * node, oh)
        return mean


class Structure:
    def __init__(self, exist, observed, dataset):
        self.exist, self.observed = np.array(exist), np.array(observed)
        self.graphical_structure = dataset.graphical_structure


class StructuredDataBatch:
    """structure.py:8-250, the methods the model and the sampler call."""

    def __init__(self, tuple_batch, dims, observed, exist, is_onehot, graphical_structure):
        self.exist = np.array(exist, dtype=np.uint8)
        self.observed = np.array([o for o, e in zip(observed, self.exist) if e], dtype=np.uint8)
        self.latent = 1 - self.observed
        self.is_onehot = [oh for oh, e in zip(is_onehot, self.exist) if e]
        self.gs = graphical_structure
        self.tuple_batch = tuple(tuple_batch)
        self._dims = dims
        self.vocab_size_features = graphical_structure.vocab_size_features
        self.name_to_index, self.names_in_batch = graphical_structure.name_to_index, graphical_structure.names_in_batch
        self.B, self.K = self.tuple_batch[0].shape[0], len(self.tuple_batch)
        assert self._dims.shape == (self.B,)

    @classmethod
    def create_copy(cls, original):
        return StructuredDataBatch(tuple(t.clone() for t in original.tuple_batch), original._dims.clone(), original.observed,
                                   original.exist, original.is_onehot, original.gs)

    def get_flat_lats(self):
        return torch.cat([t.flatten(start_dim=1) for t, o in zip(self.tuple_batch, self.observed) if not o], dim=1)

    def get_flat_lats_and_obs(self):
        return self.get_flat_lats(), tuple(t for t, o in zip(self.tuple_batch, self.observed) if o)

    def set_flat_lats(self, new_flat_lats):
        data = []
        for shape in self.gs.shapes_with_onehot():
            numel = int(np.prod(shape))
            t, new_flat_lats = new_flat_lats[:, :numel], new_flat_lats[:, numel:]
            data.append(t.reshape(-1, *shape))
        assert new_flat_lats.shape[1] == 0
        self.tuple_batch = tuple(data)

    def to(self, device):
        self.tuple_batch = tuple(t.to(device) for t in self.tuple_batch)

    def get_device(self):
        return self.tuple_batch[0].device

    def get_dims(self):
        return self._dims

    def set_dims(self, new_dims):
        self._dims = new_dims

    def delete_dims(self, new_dims):
        self.tuple_batch = tuple(self.gs.remove_problem_dims(self.tuple_batch, new_dims))
        self._dims = new_dims

    def get_mask(self, B, include_onehot_channels, include_obs):
        dev = self.get_device()
        data = [torch.ones((B, *shape), device=dev) for shape in self.gs.shapes_with_onehot()]
        data = self.gs.remove_problem_dims(data, self._dims)
        return torch.cat([t.flatten(start_dim=1) for t in data], dim=1)

    def get_next_dim_added_mask(self, B, include_onehot_channels, include_obs):
        inner = self.get_mask(B, include_onehot_channels, include_obs)
        self._dims = self._dims + 1
        outer = self.get_mask(B, include_onehot_channels, include_obs)
        self._dims = self._dims - 1
        return outer - inner


# ---- transdimensional_model.py ------------------------------------------------------------------------
class TransdimensionalEPiC(nn.Module):
    """transdimensional_model.py:135-426: EPiC trunk + rate / nearest-particle stack + vector stack."""

    def __init__(self, config, structure=None):
        super().__init__()
        self.config, self.structure = config, structure
        d, e = config.data, config.encoder
        self.max_num_particles = d.max_num_particles
        self.dim_features_continuous, self.dim_features_discrete = d.dim_features_continuous, d.dim_features_discrete
        self.vocab_size_features = d.vocab_size_features
        self.output_dim = d.dim_features_continuous + d.dim_features_discrete * d.vocab_size_features
        self.output_dim_local = e.dim_hidden_local
        self.epic = EPiCWrapper(config)
        self.add_discrete_head = e.add_discrete_head
        if self.add_discrete_head:   # constructed, never applied by forward (transdimensional_model.py:160-172 vs :265-280)
            w = d.dim_features_discrete * d.vocab_size_features
            self.fc_layer = nn.Sequential(nn.Linear(w, w), nn.SELU(), nn.Linear(w, w))
        self.noise_schedule = None
        C = self.transformer_dim = self.temb_dim = e.transformer_dim
        self.n_heads, self.n_attn_blocks = e.n_heads, e.n_attn_blocks
        self.rate_use_x0_pred = e.rate_use_x0_pred
        # False: post_rate_proj has one output, rate = softplus(.) * forward_rate(t), x0_dim_logits = 0 (model :185-188, 326-332)
        self.rdim = d.max_num_particles if self.rate_use_x0_pred else 1
        self.temb_net = nn.Linear(C, C)
        self.transformer_1_proj_in = nn.Linear(self.output_dim_local + self.vocab_size_features, C)
        self.attn_blocks = nn.ModuleList([AttnBlock(C, e.n_heads, attn_dim_reduce=1) for _ in range(e.n_attn_blocks)])
        self.res_blocks = nn.ModuleList([ResnetBlock(channels=C, dropout=0, temb_channels=C) for _ in range(e.n_attn_blocks)])
        self.pre_rate_proj = nn.Linear(C, C)
        self.post_rate_proj = nn.Linear(C, self.rdim)
        self.near_atom_proj = nn.Linear(C, 1)
        self.vec_transformer_in_proj = nn.Linear(self.output_dim_local + self.vocab_size_features + 1 + 2, C)
        self.vec_attn_blocks = nn.ModuleList([AttnBlock(C, e.n_heads, attn_dim_reduce=1) for _ in range(e.n_attn_blocks)])
        self.vec_res_blocks = nn.ModuleList([ResnetBlock(channels=C, dropout=0, temb_channels=C) for _ in range(e.n_attn_blocks)])
        self.vec_weighting_proj = nn.Linear(C, 1)
        self.pre_auto_proj = nn.Linear(C, C)
        self.post_auto_proj = nn.Linear(C, 2 * self.vocab_size_features + 1)
        self.precision = "bf16"
        self._heads_cache = None

    # ---- packing (order documented in include/mmbridge.h, mmb_trans_create) ------------------------------
    def trans_dims(self) -> "_native.TransDims":
        return _native.TransDims(self.output_dim_local, self.vocab_size_features, self.transformer_dim, self.n_heads,
                                 self.n_attn_blocks, self.max_num_particles, 0 if self.rate_use_x0_pred else 1)

    def pack_heads(self) -> torch.Tensor:
        lin = lambda m: [m.weight, m.bias]
        parts = lin(self.temb_net)
        for blocks in (self.res_blocks, self.vec_res_blocks):
            for res in blocks:
                parts += lin(res.temb_proj)

        def stack(proj_in, res_blocks, attn_blocks):
            out = lin(proj_in)
            for res, att in zip(res_blocks, attn_blocks):
                out += [res.norm1.weight, res.norm1.bias, *lin(res.conv1), res.norm2.weight, res.norm2.bias, *lin(res.conv2),
                        att.norm.weight, att.norm.bias, *lin(att.q), *lin(att.k), *lin(att.v), *lin(att.proj_out)]
            return out

        parts += stack(self.transformer_1_proj_in, self.res_blocks, self.attn_blocks)
        parts += lin(self.pre_rate_proj) + lin(self.post_rate_proj) + lin(self.near_atom_proj)
        parts += stack(self.vec_transformer_in_proj, self.vec_res_blocks, self.vec_attn_blocks)
        parts += lin(self.vec_weighting_proj) + lin(self.pre_auto_proj) + lin(self.post_auto_proj)
        return torch.cat([p.detach().to("cpu", torch.float32).reshape(-1) for p in parts])

    def native_trunk(self, device) -> "_native.EpicModel":
        return self.epic.native_model(device, None)

    def native_heads(self, device) -> "_native.TransHeads":
        params = [p for n, p in self.named_parameters() if not n.startswith(("epic.", "fc_layer."))]
        stamp = (str(device), tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        if self._heads_cache is None or self._heads_cache[0] != stamp:
            self._heads_cache = (stamp, _native.TransHeads(self.trans_dims(), self.pack_heads(), device))
        return self._heads_cache[1]

    def forward(self, st_batch, ts, nearest_atom, sample_nearest_atom=False, augment_labels=None, forward_rate=None, rnd=None,
                u_nearest=None):
        """-> (D_xt [B,N*(3+S)], rate [B,1], (auto_mean, auto_std) [B,N*(3+S)], x0_dim_logits [B,N], near_atom_logits [B,N]).
        With ``sample_nearest_atom`` the nearest particle is drawn from softmax(near_atom_logits) with one uniform per jet:
        ``u_nearest`` [B], else ``rnd.rand``, else torch.rand."""
        x, onehot = st_batch.tuple_batch[0], st_batch.tuple_batch[1]
        device = x.device
        if device.type != "cuda":
            raise _native.MmbError("TransdimensionalEPiC.forward needs CUDA tensors: libmmbridge has no CPU path")
        B = x.shape[0]
        if sample_nearest_atom:
            nearest_atom = None
            if u_nearest is None:
                u_nearest = rnd.rand((B,), device=device) if rnd is not None else torch.rand(B, device=device)
        elif nearest_atom is None:
            raise ValueError("nearest_atom is required unless sample_nearest_atom=True")
        out = _native.trans_forward(self.native_trunk(device), self.native_heads(device), x, onehot, st_batch.get_dims(), ts,
                                    nearest_atom, u_nearest, forward_rate.as_c(), precision=self.precision)
        self.last_nearest_atom = out.nearest
        return out.d_xt, out.rate.view(B, 1), (out.auto_mean, out.auto_std), out.x0_dim_logits, out.near_atom_logits


class EpsilonPrecond(nn.Module):
    """transdimensional_model.py:113-133 (``sample_nearest_atom`` / ``rnd`` passed on, which the reference omits)."""

    def __init__(self, structure, config, use_fp16=-1, **model_kwargs):
        super().__init__()
        self.structure = structure
        self.model = TransdimensionalEPiC(config, structure=structure)
        self.noise_schedule = None

    def forward(self, st_batch, ts, predict='eps', forward_rate=None, nearest_atom=None, **kw):
        eps, *others = self.model(st_batch, ts, nearest_atom=nearest_atom, forward_rate=forward_rate, **kw)
        if predict == 'eps':
            return eps, *others
        if predict == 'x0':
            xt = st_batch.get_flat_lats()
            return self.noise_schedule.predict_x0_from_xt(xt, eps, ts), *others
        raise NotImplementedError(f'predict {predict} not implemented')


class TransdimensionalJumpDiffusion(_ModuleBase):
    """transdimensional_model.py:45-111 (generation side).  ``datamodule`` may be omitted: the graphical structure
    of a ``return_type='list'`` jets batch follows from the config."""

    def __init__(self, config, datamodule=None):
        super().__init__()
        self.config = config
        gs = getattr(datamodule, "graphical_structure", None)
        if gs is None or not hasattr(gs, "max_problem_dim"):
            gs = JetsGraphicalStructure(config)
        self.graphical_structure = gs
        self.structure = Structure(gs.exist, gs.observed, SimpleNamespace(graphical_structure=gs))
        self.net = EpsilonPrecond(self.structure, config)
        lk = config.loss_kwargs
        self.forward_rate = get_forward_rate(lk.rate_function_name, config.data.max_num_particles, lk.rate_cut_t)
        self.noise_schedule = get_noise_schedule(lk.noise_schedule_name, config.data.max_num_particles, lk.vp_sde_beta_min,
                                                 lk.vp_sde_beta_max)
        # what JumpSampler.sample reads from the loss object (loss.py JumpLossFinalDim)
        self.jump_diffusion_loss = SimpleNamespace(forward_rate=self.forward_rate, noise_schedule=self.noise_schedule,
                                                   min_t=lk.min_t, loss_type=lk.loss_type, nearest_atom_pred=lk.nearest_atom_pred)

    def make_batch(self, x, onehot, dims) -> StructuredDataBatch:
        gs = self.graphical_structure
        return StructuredDataBatch((x, onehot), dims, gs.observed, gs.exist, gs.is_onehot, gs)

    def _training_not_in_scope(self, *args, **kwargs):
        raise NotImplementedError("training is outside the B200 generation hot path (SURVEY.md §8f N2)")

    training_step = validation_step = configure_optimizers = _training_not_in_scope


# ---- sampler.py ----------------------------------------------------------------------------------------
def jump_schedule(dt: float, noise_schedule: VP_SDE, no_noise_final_step: bool = False, dt_schedule: str = "uniform",
                  dt_schedule_h: float = 0.001, dt_schedule_l: float = 0.001, dt_schedule_tc: float = 0.5, corrector_steps: int = 0,
                  corrector_snr: float = 0.1, corrector_start_time: float = 0.1, corrector_finish_time: float = 0.003,
                  do_jump_corrector: bool = False, forward_rate=None) -> SimpleNamespace:
    """One row per network evaluation of JumpSampler.sample, with torch fp32 ops in the reference's order (sampler.py:79-88,
    185-231, 258-282, 315-319; noising.py:15-16, 26-39): all jets share ``ts``, so these are B-independent.

    * predictor rows (``kind`` 0) — the Euler-Maruyama step at ``ts``.  ``dt_schedule='C'`` only changes the time GRID
      (``ts -= h`` above ``tc``, ``l`` below): the reference keeps ``self.dt`` in the update coefficients, the birth
      probability and the stopping rule (sampler.py:187, 221-222, 238).
    * corrector rows (``kind`` 1) — ``corrector_steps`` Langevin evaluations at ``ts - dt`` after every predictor step with
      ``corrector_finish_time < ts < corrector_start_time`` (:199-202): ``c_score`` holds alpha = 1 - dt beta(ts - dt),
      ``c_noise`` 1/0 (noise on/off), ``death_prob`` = forward_rate(ts - dt) dt for the jump corrector (:289)."""
    if dt_schedule not in ("uniform", "C"):
        raise NotImplementedError(dt_schedule)
    ts = torch.ones((1,))
    finish_at = dt / 2

    def get_dt(ts):
        if dt_schedule == "uniform":
            return dt
        return (ts > dt_schedule_tc).long() * dt_schedule_h + (ts <= dt_schedule_tc).long() * dt_schedule_l

    def beta_of(t):
        return t * noise_schedule.beta_max + (1 - t) * noise_schedule.beta_min

    def inv_std_of(t):
        log_term = -0.25 * t ** 2 * (noise_schedule.beta_max - noise_schedule.beta_min) - 0.5 * t * noise_schedule.beta_min
        return (1 / torch.clamp(torch.sqrt(1 - torch.exp(2. * log_term)), min=0.001)).item()

    rows = []
    will_finish = False
    while True:
        if (ts - get_dt(ts)).clamp(min=finish_at / 2).max() < finish_at:
            will_finish = True
        n_corr = corrector_steps if (ts.min() < corrector_start_time and ts.max() > corrector_finish_time) else 0
        beta = beta_of(ts)
        no_noise = n_corr == 0 and no_noise_final_step and will_finish
        rows.append((0, ts.item(), (2 - torch.sqrt(1 - beta * dt)).item(), (beta * dt).item(),
                     0.0 if no_noise else torch.sqrt(beta * dt).item(), inv_std_of(ts), 0.0))
        for ci in range(n_corr):
            tm1 = ts - dt
            quiet = ci == n_corr - 1 and no_noise_final_step and will_finish
            death = 0.0
            if do_jump_corrector:
                death = float((torch.as_tensor(forward_rate.get_rate(None, tm1), dtype=torch.float32) * dt).reshape(-1)[0])
            rows.append((1, tm1.item(), 1.0, (1 - dt * beta_of(tm1)).item(), 0.0 if quiet else 1.0, inv_std_of(tm1), death))
        ts = (ts - get_dt(ts)).clamp(min=finish_at / 2)
        if ts.max() < finish_at:
            break
    cols = list(zip(*rows))
    f = [np.ascontiguousarray(c, dtype=np.float32) for c in cols[1:]]
    return SimpleNamespace(n_steps=len(rows), kind=np.ascontiguousarray(cols[0], dtype=np.uint8), ts=f[0], c_decay=f[1], c_score=f[2],
                           c_noise=f[3], inv_std=f[4], death_prob=f[5], jump_dt=float(np.float32(dt)),
                           corrector_snr=float(corrector_snr), jump_corrector=bool(do_jump_corrector and corrector_steps > 0))


class JumpSampler:
    """sampler.py:49-324.  Constructor arguments as in the reference; ``sample`` runs the whole loop on the GPU."""

    def __init__(self, structure, dt, corrector_steps, corrector_snr, corrector_start_time, corrector_finish_time,
                 do_conditioning, condition_type, condition_sweep_idx, condition_sweep_path, guidance_weight, do_jump_corrector,
                 sample_near_atom, dt_schedule, dt_schedule_h, dt_schedule_l, dt_schedule_tc, no_noise_final_step):
        self.structure, self.dt = structure, dt
        if do_conditioning or dt_schedule not in ('uniform', 'C') or not sample_near_atom:
            # sample_near_atom=False cannot run in the reference (sampler.py:95 calls the network without a nearest particle,
            # transdimensional_model.py:341 then fails on None); conditioning asks the dataset for `condition_state`, which
            # only the vendored, unimportable QM9 dataset defines (qm9.py:1981), and differentiates through the network
            raise NotImplementedError("native JumpSampler: sample_near_atom=True, uniform or 'C' time grid, no conditioning "
                                      "(guidance needs the network's backward pass and a dataset with condition_state, which "
                                      "no particle-cloud dataset of the reference has; SURVEY.md §8f N4)")
        self.corrector_snr, self.corrector_start_time, self.corrector_finish_time = corrector_snr, corrector_start_time, corrector_finish_time
        self.do_jump_corrector = do_jump_corrector
        self.dt_schedule, self.dt_schedule_h, self.dt_schedule_l, self.dt_schedule_tc = dt_schedule, dt_schedule_h, dt_schedule_l, dt_schedule_tc
        self.corrector_steps, self.sample_near_atom, self.no_noise_final_step = corrector_steps, sample_near_atom, no_noise_final_step
        self.seed = 0
        self._jets_generated = 0

    def get_dt(self, ts):
        if self.dt_schedule == 'uniform':
            return self.dt
        return (ts > self.dt_schedule_tc).long() * self.dt_schedule_h + (ts <= self.dt_schedule_tc).long() * self.dt_schedule_l

    @torch.no_grad()
    def sample(self, net, in_st_batch, loss, rnd=None, known_dims=None, dataset_obj=None, noise=None, jet_offset=None,
               precision=None):
        """-> StructuredDataBatch with the generated jets (on the GPU).  ``noise``: optional namespace of injected draws
        (z_init [B,N*F], z_diff [n_steps,B,N*F], u_near [n_steps,B], u_jump [n_steps,B], z_new [n_steps,B,F])."""
        net.noise_schedule = net.model.noise_schedule = loss.noise_schedule
        model = net.model
        state = StructuredDataBatch.create_copy(in_st_batch)
        device = state.get_device()
        if device.type != "cuda":
            if not torch.cuda.is_available():
                raise _native.MmbError("JumpSampler.sample needs a CUDA device: libmmbridge has no CPU path")
            device = torch.device("cuda", torch.cuda.current_device())
            state.to(device)
        x0 = state.get_flat_lats()
        B = x0.shape[0]
        if jet_offset is None:
            jet_offset = next_jet_offset(self, B)
        # x_T ~ N(0, I), one particle per jet, centred (sampler.py:170-183)
        if noise is not None and getattr(noise, "z_init", None) is not None:
            xT = noise.z_init.to(device, torch.float32)
        elif rnd is not None:
            xT = rnd.randn_like(x0)
        else:
            g = torch.Generator(device=device).manual_seed(self.seed * 1000003 + jet_offset)
            xT = torch.randn(x0.shape, device=device, generator=g)
        state.set_flat_lats(xT)
        dims = torch.ones((B,), dtype=torch.int32, device=device)
        state.delete_dims(new_dims=dims)
        state.gs.adjust_st_batch(state)
        sched = jump_schedule(self.dt, loss.noise_schedule, self.no_noise_final_step, self.dt_schedule, self.dt_schedule_h,
                              self.dt_schedule_l, self.dt_schedule_tc, self.corrector_steps, self.corrector_snr,
                              self.corrector_start_time, self.corrector_finish_time, self.do_jump_corrector, loss.forward_rate)
        self.last_n_steps = sched.n_steps   # the reference prints this as 'nfe' (sampler.py:323)
        x, onehot = state.tuple_batch[0].contiguous(), state.tuple_batch[1].contiguous()
        _native.trans_sample(model.native_trunk(device), model.native_heads(device), x, onehot, dims, sched,
                             loss.forward_rate.as_c(), noise=noise, seed=self.seed, jet_offset=jet_offset,
                             precision=precision or model.precision)
        state.tuple_batch = (x, onehot)
        state.set_dims(dims.to(torch.int64))
        return state
