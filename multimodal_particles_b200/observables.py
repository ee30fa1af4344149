"""Post-processing and jet-level observables of generated jets — drop-in for the generation-side use of
``ParticleClouds`` (mp/data/particle_clouds/particles.py:22-156: construction from a ``HybridState``, ``postprocess``,
``compute_4mom``) and ``JetClassHighLevelFeatures`` (mp/data/particle_clouds/jets.py:85-141, 329-332: jet kinematics,
multiplicity, jet charges, 1-D Wasserstein distance), SURVEY.md §8f N1.  One fused kernel (``mmb_jet_observables``) does the
de-standardisation, ``tokens_to_physics``, the 4-momenta and the per-jet sums; only the clustering-based substructure
(fastjet N-subjettiness / D2, jets.py:204-240) is not provided.
"""
import ctypes

import numpy as np
import torch

from . import _native
from .epic import as_u8

JET_COLUMNS = ("px", "py", "pz", "e", "pt", "m", "eta", "phi", "multiplicity", "Q_total", "Q_jet")


def jet_observables(x, k_u8, mask_u8, stats=None, want_particles=True):
    """x [B,N,3] f32, k/mask [B,N] u8 on the GPU -> (x_phys [B,N,3] | None, flavor_charge [B,N,2] int8 | None, jets [B,11])."""
    _native._require_cuda(x, k_u8, mask_u8)
    B, N, _ = x.shape
    dev = x.device
    x_phys = torch.empty_like(x) if want_particles else None
    fc = torch.empty(B, N, 2, dtype=torch.int8, device=dev) if want_particles else None
    jets = torch.empty(B, len(JET_COLUMNS), dtype=torch.float32, device=dev)
    mean = std = None
    if stats is not None:
        mean = (ctypes.c_float * 3)(*[float(v) for v in stats["mean"][:3]])
        std = (ctypes.c_float * 3)(*[float(v) for v in stats["std"][:3]])
    with torch.cuda.device(dev):
        _native.check(_native.load().mmb_jet_observables(_native._ptr(x), _native._ptr(k_u8), _native._ptr(mask_u8), mean, std, B, N,
                                                         _native._ptr(x_phys), _native._ptr(fc), _native._ptr(jets), _native._stream()))
    return x_phys, fc, jets


class ParticleClouds:
    """Generated jets as the reference's container: ``ParticleClouds(dataset=HybridState)`` (particles.py:33-38)."""

    def __init__(self, dataset, data_paths=None, **data_params):
        if not hasattr(dataset, "continuous"):
            raise NotImplementedError("only construction from a generated HybridState is part of the generation path")
        mask = getattr(dataset, "absorbing", None)
        if mask is None:
            mask = getattr(dataset, "mask_t")
        self.continuous, self.discrete, self.mask = dataset.continuous, dataset.discrete, mask
        self._set_views()
        self.multiplicity = torch.sum(self.mask, dim=1)
        self._jets = None

    def _set_views(self):
        self.pt, self.eta_rel, self.phi_rel = self.continuous[..., 0], self.continuous[..., 1], self.continuous[..., 2]

    def __len__(self):
        return self.continuous.shape[0]

    def _run(self, stats):
        dev = self.continuous.device
        if dev.type != "cuda":
            if not torch.cuda.is_available():
                raise _native.MmbError("post-processing needs a CUDA device: libmmbridge has no CPU path")
            dev = torch.device("cuda", torch.cuda.current_device())
        x = self.continuous.to(dev, torch.float32).contiguous()
        return jet_observables(x, as_u8(self.discrete.to(dev)), as_u8(self.mask.to(dev)), stats)

    def postprocess(self, input_continuous="standardize", input_discrete="tokens", stats=None):
        """In place, like the reference: continuous de-standardised and masked; ``flavor`` one-hot [B,N,5], ``charge`` [B,N,1]
        and ``discrete`` = cat(flavor, charge), all masked.  The jet sums computed in the same pass are kept for
        ``JetClassHighLevelFeatures``."""
        if input_discrete != "tokens":
            raise NotImplementedError("native post-processing covers input_discrete='tokens' (every shipped config)")
        stats = getattr(self, "stats", stats)
        out_dev = self.continuous.device
        x_phys, fc, jets = self._run(stats if input_continuous == "standardize" else None)
        live = as_u8(self.mask.to(fc.device)).bool()
        flavor = torch.nn.functional.one_hot(fc[..., 0].long(), num_classes=5) * live[..., None]
        self.continuous = x_phys.to(out_dev)
        self.flavor, self.charge = flavor.to(out_dev), fc[..., 1:2].long().to(out_dev)
        self.discrete = torch.cat([self.flavor, self.charge], dim=-1)
        self._set_views()
        self._jets = jets.to(out_dev)

    def compute_4mom(self):
        self.px = self.pt * torch.cos(self.phi_rel)
        self.py = self.pt * torch.sin(self.phi_rel)
        self.pz = self.pt * torch.sinh(self.eta_rel)
        self.e = self.pt * torch.cosh(self.eta_rel)


class JetClassHighLevelFeatures:
    """Jet kinematics, multiplicity and charges of post-processed constituents (jets.py:85-107, 138-141)."""

    def __init__(self, constituents: ParticleClouds):
        self.constituents = constituents
        jets = constituents._jets
        if jets is None:   # constituents already in physical units (no postprocess call): one pass without de-standardisation
            c = constituents
            dev = c.continuous.device if c.continuous.is_cuda else torch.device("cuda", torch.cuda.current_device())
            tokens = c.discrete if c.discrete.shape[-1] == 1 else None
            if tokens is None:
                raise NotImplementedError("construct from token-valued clouds or call postprocess() first")
            _, _, jets = jet_observables(c.continuous.to(dev, torch.float32).contiguous(), as_u8(tokens.to(dev)), as_u8(c.mask.to(dev)),
                                         None, want_particles=False)
            jets = jets.to(c.continuous.device)
        for i, name in enumerate(JET_COLUMNS):
            setattr(self, name, jets[:, i])
        self.multiplicity = jets[:, JET_COLUMNS.index("multiplicity")].long().unsqueeze(-1)

    def substructure(self):
        raise NotImplementedError("fastjet substructure (jets.py:204-240) stays on the CPU side of the reference")

    def Wassertein1D(self, feature, reference):
        import scipy.stats
        x, y = getattr(self, feature), getattr(reference, feature)
        return scipy.stats.wasserstein_distance(np.asarray(x.detach().cpu()).reshape(-1), np.asarray(y.detach().cpu()).reshape(-1))
