"""Solver steps of the hybrid bridge: thin host wrappers over the fused update kernel.

Interface as in the reference (mp/models/generative/bridges.py): ``bridge.solver_step(state,
heads, delta_t[, multimodal])`` mutates ``state`` and returns it.  A lone ``solver_step`` runs the
fused kernel with the other sub-steps switched off (MMB_FLAG_NO_EULER / MMB_FLAG_NO_JUMP);
``simulate_dynamics`` never calls these one by one — it uses the whole fused update.

Differences from the reference, by design:

* the discrete step draws ONE uniform per particle and applies the exactly equivalent categorical
  form of the reference's gated Poisson tau-leap (SURVEY.md §A.4 Form B), the
  absorbing step ``u < p`` instead of ``torch.bernoulli``; pass ``uniforms=`` to inject the draws
  (parity tests), otherwise ``torch.rand`` on the state's device;
* tokens/masks are narrowed to uint8 for the kernel and written back in the state's int64 layout.

The training-time samplers of the continuous and discrete bridges (bridges.py:23-33,99-104,134-177) run fused inside
``MultiModalBridgeMatching.sample_bridges`` (``mmb_sample_bridges``); ``AbsorbingBridge.sample`` (bridges.py:233-249) is
mirrored here.
"""
import torch

from . import _native
from .epic import as_u8
from .sharding import next_jet_offset
from .steptable import survival_probability, telegraph_coefficients

NO_EULER, NO_JUMP = 2, 4


def _shared_time(state) -> torch.Tensor:
    """Generation feeds one time for the whole batch (mbm.py:211) and the kernel takes the step's
    coefficients as scalars; a batch with differing times is rejected rather than mis-stepped."""
    t = state.time.reshape(-1).float().cpu()
    if t.numel() > 1 and not bool((t == t[0]).all()):
        raise NotImplementedError("solver_step on the native path needs one shared time per batch")
    return t[:1]


def _mask_of(state, heads, multimodal):
    return heads.absorbing if multimodal else state.mask_t


class LinearUniformBridge:
    """Conditional-OT flow matching; Euler step ``x <- (x + dt v) * mask`` (bridges.py:38-45)."""

    def __init__(self, config):
        self.sigma = config.bridge.sigma

    def solver_step(self, state, heads, delta_t, multimodal: bool = True):
        x = state.continuous
        mask = as_u8(_mask_of(state, heads, multimodal))
        k_dummy = torch.zeros_like(mask)
        _native.bridge_update(x, k_dummy, mask, heads.continuous.contiguous(), None, None,
                              float(delta_t), 0.0, 0.0, flags=NO_JUMP)
        return state


class TelegraphBridge:
    """Multivariate telegraph bridge on tokens (bridges.py:86-201)."""

    def __init__(self, config):
        self.gamma = config.bridge.gamma
        self.time_epsilon = config.bridge.time_eps
        self.vocab_size = config.data.vocab_size_features

    def solver_step(self, state, heads, delta_t, multimodal: bool = True, uniforms=None):
        k64 = state.discrete
        assert bool((k64 >= 0).all()) and bool((k64 < self.vocab_size).all()), \
            "Values in `k` outside of bound! k_min={}, k_max={}".format(k64.min(), k64.max())
        bc, cc = telegraph_coefficients(_shared_time(state), self.vocab_size, self.gamma)
        B, N = k64.shape[0], k64.shape[1]
        dev = k64.device
        u = torch.rand(B, N, device=dev) if uniforms is None else uniforms.reshape(B, N).to(dev).contiguous()
        k = as_u8(k64)
        _native.bridge_update(state.continuous, k, as_u8(_mask_of(state, heads, multimodal)), None,
                              heads.discrete.contiguous(), u, float(delta_t), float(bc), float(cc), flags=NO_EULER)
        state.discrete = k.to(k64.dtype).unsqueeze(-1)
        return state


class AbsorbingBridge:
    """Birth of particles: ``mask' = mask | (u < min(1, dt SP(t) sigmoid(a)))`` (bridges.py:203-286)."""

    def __init__(self, config):
        self.gamma_absorb = torch.tensor(config.bridge.gamma_absorb, dtype=torch.float32)
        self.time_epsilon = config.bridge.time_eps
        self.vocab_size = 2
        self.seed = 0   # Philox key of sample() when no uniforms are injected (AbsorbingFlow.sample_bridges sets its own)

    def survival_probability(self, t):
        return survival_probability(t, float(self.gamma_absorb))

    def sample(self, time, target_mask, uniforms=None):
        """time [B,1,1], target_mask [B,N,1] -> mask_t [B,N,1] int64: particles alive at ``time`` (bridges.py:233-249);
        ``uniforms`` [B,N] injects the draws, default in-kernel Philox."""
        dev = target_mask.device
        B, N = target_mask.shape[0], target_mask.shape[1]
        sp = self.survival_probability(time.reshape(B).float().cpu()).to(dev).contiguous()
        u = None if uniforms is None else uniforms.reshape(B, N).to(dev, torch.float32).contiguous()
        return _native.absorbing_sample(sp, as_u8(target_mask), u, seed=self.seed,
                                        jet_offset=next_jet_offset(self, B, "_sampled")).to(torch.int64).unsqueeze(-1)

    def solver_step(self, state, heads, delta_t, uniforms=None):
        m64 = state.mask_t
        B, N = m64.shape[0], m64.shape[1]
        dev = m64.device
        sp = self.survival_probability(_shared_time(state))
        u = torch.rand(B, N, device=dev) if uniforms is None else uniforms.reshape(B, N).to(dev).contiguous()
        mask = as_u8(m64)
        k_dummy = torch.zeros_like(mask)
        _native.bridge_update(state.continuous, k_dummy, mask, None, None, None, float(delta_t), 0.0, 0.0,
                              absorb_logit=heads.absorbing.reshape(B, N).contiguous(), u_absorb=u, sp=float(sp),
                              flags=_native.FLAG_ABSORBING | NO_EULER | NO_JUMP)
        state.mask_t = mask.to(torch.int64).unsqueeze(-1)
        return state
