"""State and head containers of the generation loop.

Same attribute names, shapes and dtypes as the reference (mp/models/generative/
multimodal_bridge_matching.py:13-75, mp/models/generative/absorbing/states.py:8-71):
``time [B,1]`` f32 at generation time, ``continuous [B,N,Dc]`` f32, ``discrete [B,N,1]`` int64,
``absorbing`` / ``mask_t [B,N,1]`` int64.
"""
from dataclasses import dataclass, fields
from typing import List, Optional

import torch


class _TensorRecord:
    """Field-wise tensor transforms shared by both state types."""

    def _map(self, fn):
        values = {}
        for f in fields(self):
            value = getattr(self, f.name)
            values[f.name] = fn(value) if isinstance(value, torch.Tensor) else None
        return type(self)(**values)

    def to(self, device):
        return self._map(lambda t: t.to(device))

    def cpu(self):
        return self._map(lambda t: t.cpu())

    def clone(self):
        return self._map(lambda t: t.clone())

    @property
    def device(self):
        for f in fields(self):
            value = getattr(self, f.name)
            if isinstance(value, torch.Tensor) and f.name != "time":
                return value.device
        return torch.device("cpu")

    @classmethod
    def _cat(cls, states, dim, sources):
        out = {}
        for name, source in sources.items():
            parts = [getattr(s, source, None) for s in states]
            parts = [p for p in parts if p is not None]
            out[name] = torch.cat(parts, dim=dim) if parts else None
        return cls(**out)


@dataclass
class HybridState(_TensorRecord):
    """time-dependent hybrid bridge state (t, x, k, mask)   [mbm.py:13-69]"""

    time: Optional[torch.Tensor] = None
    continuous: Optional[torch.Tensor] = None
    discrete: Optional[torch.Tensor] = None
    absorbing: Optional[torch.Tensor] = None

    def detach(self):
        return self._map(lambda t: t.detach())

    @staticmethod
    def cat(states: List["HybridState"], dim=0) -> "HybridState":
        names = ("time", "continuous", "discrete", "absorbing")
        return HybridState._cat(states, dim, {n: n for n in names})


@dataclass
class MultiHeadOutput:
    """[mbm.py:71-75]"""

    continuous: Optional[torch.Tensor] = None
    discrete: Optional[torch.Tensor] = None
    absorbing: Optional[torch.Tensor] = None


@dataclass
class OutputHeads:
    """[absorbing/states.py:8-12]"""

    continuous: Optional[torch.Tensor] = None
    discrete: Optional[torch.Tensor] = None
    absorbing: Optional[torch.Tensor] = None


@dataclass
class AbsorbingBridgeState(_TensorRecord):
    """[absorbing/states.py:15-71]; ``detach`` acts in place and returns self as the reference's does."""

    time: Optional[torch.Tensor] = None
    continuous: Optional[torch.Tensor] = None
    discrete: Optional[torch.Tensor] = None
    mask_t: Optional[torch.Tensor] = None

    def detach(self):
        for f in fields(self):
            value = getattr(self, f.name)
            if value is not None:
                setattr(self, f.name, value.detach())
        return self

    @staticmethod
    def cat(states: List["AbsorbingBridgeState"], dim=0) -> "AbsorbingBridgeState":
        # the reference reads the attribute "absorbing" for mask_t (states.py:54), which no
        # AbsorbingBridgeState has, so mask_t of a concatenation is always None; kept.
        return AbsorbingBridgeState._cat(
            states, dim, {"time": "time", "continuous": "continuous", "discrete": "discrete", "mask_t": "absorbing"})
