"""AbsorbingFlow — drop-in for the generation side of
mp/models/generative/absorbing/absorbing_flows.py (:14-159 generator, :161-275 flow).

Same constructor argument (``AbsorbingConfig``), ``forward(state, batch) -> OutputHeads``,
``simulate_dynamics(state, batch) -> AbsorbingBridgeState`` (CPU, detached), state-dict keys
(``generator.epic.*``, ``generator.discrete_head_mlp.*``, ``generator.temb_net``,
``generator.transformer_1_proj_in``, ``generator.{attn,res}_blocks.*``, ``generator.{pre,post}_rate_proj``,
``loss_multihead.weights``).  Per step: EPiC trunk (with the last local hidden), discrete MLP head,
the transformer absorbing-rate head, then the fused update in the reference's order — birth of
particles first, Euler and token jump with the NEW mask (absorbing_flows.py:271-273).
``predict_step`` of the reference is broken (reads ``config.pipeline``) and is not mirrored.
"""
import math

import torch
from torch import nn

from . import _native
from .bridges import AbsorbingBridge, LinearUniformBridge, TelegraphBridge
from .epic import EPiCWrapper, as_u8
from .multimodal_bridge_matching import MultiHeadLoss, _ModuleBase, to_host_async
from .states import AbsorbingBridgeState, OutputHeads
from .sharding import next_jet_offset
from .steptable import build_step_table


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, max_timesteps: int = 10000) -> torch.Tensor:
    """[sin(t f), cos(t f)], f_j = exp(-j ln(max)/(half-1))   (mp/models/architectures/gsdm.py:8-26)"""
    half = embedding_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_timesteps) / (half - 1)))
    emb = timesteps.float()[:, None] * freq.to(timesteps.device)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


def _conv(a, b):
    return nn.Conv1d(a, b, kernel_size=1, stride=1, padding=0)


def _norm(channels):
    return nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6, affine=True)


class AttnBlock(nn.Module):
    """Parameter container of gsdm.AttnBlock (gsdm.py:69-95)."""

    def __init__(self, in_channels, n_heads=1, attn_dim_reduce=1):
        super().__init__()
        self.in_channels, self.n_heads = in_channels, n_heads
        self.norm = _norm(in_channels)
        self.q = _conv(in_channels, in_channels // attn_dim_reduce)
        self.k = _conv(in_channels, in_channels // attn_dim_reduce)
        self.v = _conv(in_channels, in_channels // attn_dim_reduce)
        self.proj_out = _conv(in_channels // attn_dim_reduce, in_channels)


class ResnetBlock(nn.Module):
    """Parameter container of gsdm.ResnetBlock (gsdm.py:38-52)."""

    def __init__(self, *, channels, dropout=0, temb_channels=512):
        super().__init__()
        self.norm1 = _norm(channels)
        self.conv1 = _conv(channels, channels)
        self.temb_proj = _conv(temb_channels, channels)
        self.norm2 = _norm(channels)
        self.conv2 = _conv(channels, channels)


class AbsorbingGenerator(nn.Module):
    """EPiC trunk + discrete MLP head + transformer absorbing-rate head (absorbing_flows.py:14-159)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        d, g = config.data, config.generator
        self.max_num_particles = d.max_num_particles
        self.dim_features_continuous = d.dim_features_continuous
        self.dim_features_discrete = d.dim_features_discrete
        self.vocab_size_features = d.vocab_size_features
        self.encoder_output_dim = d.dim_features_continuous + d.dim_features_discrete * d.vocab_size_features
        self.encoder_output_dim_local = config.encoder.dim_hidden_local
        self.epic = EPiCWrapper(config)
        if self.epic.embedding.dim_context:
            # absorbing_flows.py:150 reads the context with a trailing comma (a one-element tuple), EPiCWrapper.forward turns the
            # non-tensor into None and the context embedding fails: the reference has no behaviour with context to match
            raise NotImplementedError("AbsorbingGenerator with context features: the reference's forward fails there "
                                      "(absorbing_flows.py:150 wraps batch.context_continuous in a tuple); not built")
        self.add_discrete_head = config.encoder.add_discrete_head
        if self.add_discrete_head:
            width = d.dim_features_discrete * d.vocab_size_features
            self.discrete_head_mlp = nn.Sequential(nn.Linear(width, g.discrete_head_hidden_dim), nn.SELU(),
                                                   nn.Linear(g.discrete_head_hidden_dim, width))
        self.transformer_dim = self.temb_dim = g.transformer_dim
        self.n_heads, self.n_attn_blocks = g.n_heads, g.n_attn_blocks
        self.temb_net = nn.Linear(self.temb_dim, self.temb_dim)
        self.transformer_1_proj_in = nn.Linear(self.encoder_output_dim_local + 2, self.transformer_dim)
        self.attn_blocks = nn.ModuleList([AttnBlock(self.transformer_dim, g.n_heads, attn_dim_reduce=1)
                                          for _ in range(g.n_attn_blocks)])
        self.res_blocks = nn.ModuleList([ResnetBlock(channels=self.transformer_dim, dropout=0, temb_channels=self.temb_dim)
                                         for _ in range(g.n_attn_blocks)])
        self.pre_rate_proj = nn.Linear(self.transformer_dim, self.transformer_dim)
        self.post_rate_proj = nn.Linear(self.transformer_dim, 1)
        self.precision = "bf16"
        self._head_cache = None

    # ---- packing (order documented in include/mmbridge.h, mmb_absorb_head_create) ------------------
    def pack_head_weights(self) -> torch.Tensor:
        parts = [self.transformer_1_proj_in.weight, self.transformer_1_proj_in.bias]
        for res, att in zip(self.res_blocks, self.attn_blocks):
            parts += [res.norm1.weight, res.norm1.bias, res.conv1.weight, res.conv1.bias,
                      res.norm2.weight, res.norm2.bias, res.conv2.weight, res.conv2.bias,
                      att.norm.weight, att.norm.bias]
            for conv in (att.q, att.k, att.v, att.proj_out):
                parts += [conv.weight, conv.bias]
        parts += [self.pre_rate_proj.weight, self.pre_rate_proj.bias, self.post_rate_proj.weight, self.post_rate_proj.bias]
        return torch.cat([p.detach().to("cpu", torch.float32).reshape(-1) for p in parts])

    def time_bias(self, t: torch.Tensor) -> torch.Tensor:
        """t [M] f32 (CPU) -> [M, n_blocks, C]: temb_proj_b(swish(temb_net(timestep_embedding(1000 t)))), i.e. the
        ``h + temb_proj(nonlinearity(temb))`` term of every ResnetBlock (gsdm.py:58, absorbing_flows.py:108-111),
        with torch fp32 ops on the host (per-step constants of the generation loop)."""
        dev = self.temb_net.weight.device
        temb = self.temb_net(get_timestep_embedding(t.to(dev) * 1000, self.temb_dim))
        act = temb * torch.sigmoid(temb)
        out = [blk.temb_proj(act[:, :, None])[:, :, 0] for blk in self.res_blocks]
        return torch.stack(out, 1).detach().float().contiguous()

    def native_trunk(self, device) -> _native.EpicModel:
        return self.epic.native_model(device, self.discrete_head_mlp if self.add_discrete_head else None)

    def native_head(self, device) -> "_native.AbsorbHead":
        params = [p for n, p in self.named_parameters() if not n.startswith(("epic.", "discrete_head_mlp."))]
        stamp = (str(device), tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        if self._head_cache is None or self._head_cache[0] != stamp:
            head = _native.AbsorbHead(self.encoder_output_dim_local, self.transformer_dim, self.n_heads,
                                      self.n_attn_blocks, self.pack_head_weights(), device)
            self._head_cache = (stamp, head)
        return self._head_cache[1]

    def forward(self, state: AbsorbingBridgeState, batch=None) -> OutputHeads:
        x = state.continuous.contiguous().float()
        dev = x.device
        mask = as_u8(state.mask_t)
        if not bool(mask.any()):
            # one_hot(mask_t.squeeze()) infers a single class there and the reference crashes
            raise RuntimeError("absorbing head needs at least one live particle in the batch (reference: one_hot width 1)")
        t = state.time.detach().reshape(state.time.shape[0]).float().cpu()
        temb = self.epic.time_embedding(state.time).to(dev)
        v, logits, hidden = self.native_trunk(dev).forward(x, as_u8(state.discrete), mask, temb, want_hidden=True,
                                                           precision=self.precision)
        rate_logit = self.native_head(dev).forward(hidden, mask, self.time_bias(t).to(dev))
        return OutputHeads(v, logits, rate_logit.unsqueeze(-1))


class AbsorbingFlow(_ModuleBase):
    """Model for hybrid data with varying size (absorbing_flows.py:161-275), generation side."""

    def __init__(self, config, precision: str = "bf16"):
        super().__init__()
        self.config = config
        self.vocab_size = config.data.vocab_size_features
        self.generator = AbsorbingGenerator(config)
        self.generator.precision = precision
        self.bridge_continuous = LinearUniformBridge(config)
        self.bridge_discrete = TelegraphBridge(config)
        self.bridge_absorbing = AbsorbingBridge(config)
        self.loss_multihead = MultiHeadLoss(number_of_losses=3)
        self.min_t = config.bridge.time_eps
        self.precision = precision
        self.seed = 0
        self._jets_generated = 0
        self.save_hyperparameters()

    def forward(self, state: AbsorbingBridgeState, batch=None) -> OutputHeads:
        return self.generator(state, batch)

    def step_table(self):
        b, e = self.config.bridge, self.config.encoder
        return build_step_table(b.num_timesteps, b.time_eps, self.vocab_size, b.gamma, e.dim_emb_time,
                                gamma_absorb=b.gamma_absorb)

    @torch.no_grad()
    def simulate_dynamics(self, state: AbsorbingBridgeState, batch=None, uniforms_jump=None, uniforms_absorb=None,
                          precision=None, jet_offset=None, return_device=False) -> AbsorbingBridgeState:
        """Generate target data from the source state (absorbing_flows.py:255-275): per step the three heads
        from the OLD mask, then birth -> Euler -> jump with the new mask.  ``uniforms_*`` [T-1,B,N] inject the
        draws (parity); default: in-kernel Philox streams 0 (jump) and 1 (birth)."""
        gen = self.generator
        device = self.device if self.device.type == "cuda" else state.continuous.device
        if device.type != "cuda":
            if not torch.cuda.is_available():
                raise _native.MmbError("generation needs a CUDA device: libmmbridge has no CPU path")
            device = torch.device("cuda", torch.cuda.current_device())
        table = self.step_table()
        x = state.continuous.to(device, torch.float32, copy=True).contiguous()
        k64 = state.discrete
        assert bool((k64 >= 0).all()) and bool((k64 < self.vocab_size).all()), "Values in `k` outside of bound!"
        k, mask = as_u8(k64.to(device)), as_u8(state.mask_t.to(device)).clone()
        B, N, _ = x.shape
        if jet_offset is None:
            jet_offset = next_jet_offset(self, B)
        prep = lambda u: None if u is None else u.to(device, torch.float32).reshape(table.n_steps, B, N).contiguous()
        _native.generate_absorbing(gen.native_trunk(device), gen.native_head(device), x, k, mask, table,
                                   gen.time_bias(table.t), prep(uniforms_jump), prep(uniforms_absorb),
                                   seed=self.seed, jet_offset=jet_offset, precision=precision or self.precision)
        if return_device:
            return AbsorbingBridgeState(time=torch.full((B, 1), float(table.t[-1]), device=device), continuous=x,
                                        discrete=k.to(k64.dtype).unsqueeze(-1), mask_t=mask.to(torch.int64).unsqueeze(-1))
        # compact tensors cross PCIe into page-locked memory; tokens and masks are widened to int64 on the host
        x_host, k_host, m_host = to_host_async(x), to_host_async(k), to_host_async(mask)
        torch.cuda.current_stream(device).synchronize()
        return AbsorbingBridgeState(time=torch.full((B, 1), float(table.t[-1])), continuous=x_host,
                                    discrete=k_host.to(k64.dtype).unsqueeze(-1), mask_t=m_host.to(torch.int64).unsqueeze(-1))

    def reshape_time(self, t, x):
        return t if isinstance(t, (float, int)) else t.reshape(-1, *([1] * (x.dim() - 1)))

    @torch.no_grad()
    def sample_bridges(self, batch, t=None, z=None, u=None, u_absorb=None) -> AbsorbingBridgeState:
        """Sample stochastic bridges (absorbing_flows.py:187-208): ``t = min_t + (1 - min_t) U``, the continuous and telegraph
        bridges in one kernel (``mmb_sample_bridges``) and the absorbing bridge's mask (``mmb_absorbing_sample``).
        ``t`` [B], ``z`` [B,N,3], ``u`` / ``u_absorb`` [B,N] inject the draws."""
        device = batch.target_continuous.device
        if device.type != "cuda":
            if not torch.cuda.is_available():
                raise _native.MmbError("sample_bridges needs a CUDA device: libmmbridge has no CPU path")
            device = torch.device("cuda", torch.cuda.current_device())
        x1 = batch.target_continuous.to(device, torch.float32).contiguous()
        x0 = batch.source_continuous.to(device, torch.float32).contiguous()
        B = x1.shape[0]
        if t is None:
            t = self.min_t + (1 - self.min_t) * torch.rand(B, device=device)
        t = t.to(device, torch.float32).contiguous()
        prep = lambda a: None if a is None else a.to(device, torch.float32).contiguous()
        off = next_jet_offset(self, B, "_bridges_sampled")
        xt, kt = _native.sample_bridges(x0, x1, as_u8(batch.source_discrete.to(device)), as_u8(batch.target_discrete.to(device)), t,
                                        self.bridge_continuous.sigma, self.bridge_discrete.gamma, self.vocab_size, prep(z), prep(u),
                                        seed=self.seed, jet_offset=off)
        time = self.reshape_time(t, x1)
        self.bridge_absorbing.seed = self.seed
        mask_t = self.bridge_absorbing.sample(time, batch.target_mask.to(device), uniforms=u_absorb)
        return AbsorbingBridgeState(time, xt, kt.long().unsqueeze(-1), mask_t)

    def _training_not_in_scope(self, *args, **kwargs):
        raise NotImplementedError("losses / backward of the absorbing flow are outside the B200 hot path (SURVEY.md §8f N2: "
                                  "bridge sampling only)")

    loss_continuous = loss_discrete = loss_absorbing = _training_not_in_scope
    training_step = validation_step = configure_optimizers = _training_not_in_scope
