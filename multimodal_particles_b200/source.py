"""Source-state construction on the GPU — the generation-side use of ``sample_noise("GaussNoise")`` and
``sample_masks`` (mp/data/particle_clouds/utils.py:222-286) plus the token conversion of ``ParticleClouds.preprocess``
(particles.py:111-113), as one kernel (``mmb_sample_source``) with Philox draws keyed by the global jet index."""
import ctypes

import numpy as np
import torch

from . import _native
from .states import HybridState


def multiplicity_cdf(target_multiplicity, max_num_particles: int) -> np.ndarray:
    """Cumulative multiplicity probabilities exactly as sample_masks builds them (utils.py:266-277): density histogram of
    the target multiplicities over the unit bins 0..max_num_particles, normalised in fp32."""
    hist_values, _ = np.histogram(np.asarray(target_multiplicity), bins=np.arange(0, max_num_particles + 2, 1), density=True)
    h = torch.tensor(hist_values, dtype=torch.float)
    probs = h / h.sum()
    return torch.cumsum(probs, 0).numpy().astype(np.float32)


def sample_source_state(num_jets: int, max_num_particles: int = 128, target_multiplicity=None, min_num_particles: int = 0,
                        scale: float = 1.0, cat_probs=(0.2, 0.2, 0.2, 0.2, 0.2), device="cuda", seed: int = 0, jet_offset: int = 0,
                        compact: bool = False, out=None):
    """-> HybridState(None, continuous [B,N,3] f32, discrete [B,N,1] int64, absorbing = mask [B,N,1] int64) on ``device``
    (``compact=True``: the uint8 [B,N] tensors the kernels consume, no widening; ``out=(x, k, mask)``: write into existing
    device tensors, e.g. the views of a ``sharding.PackedJets``)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _native.MmbError("sample_source_state needs a CUDA device: libmmbridge has no CPU path")
    B, N = num_jets, max_num_particles
    if out is not None:
        x, k, mask = out
        _native._require_cuda(x, k, mask)
        assert x.shape == (B, N, 3) and k.shape == (B, N) and mask.shape == (B, N)
    else:
        x = torch.empty(B, N, 3, device=device)
        k = torch.empty(B, N, dtype=torch.uint8, device=device)
        mask = torch.empty(B, N, dtype=torch.uint8, device=device)
    cdf = None
    if target_multiplicity is not None and min_num_particles != max_num_particles:   # sample_masks' two "all ones" exits
        cdf = torch.from_numpy(multiplicity_cdf(target_multiplicity, N)).to(device)
    probs = (ctypes.c_float * 5)(*[float(p) for p in cat_probs])
    with torch.cuda.device(device):
        _native.check(_native.load().mmb_sample_source(_native._ptr(x), _native._ptr(k), _native._ptr(mask), B, N, float(scale), probs,
                                                       _native._ptr(cdf), seed, jet_offset, _native._stream()))
    if compact:
        return x, k, mask
    return HybridState(None, x, k.long().unsqueeze(-1), mask.long().unsqueeze(-1))
