"""Per-step scalars of the generation loop, computed on the host with torch fp32 ops in the
reference's own order so that they equal what the reference feeds its kernels bit for bit
(SURVEY.md §A.1, §A.4).  One row per solver step i = 1..T-1:

    t     = linspace(0, 1-eps, T)[i]                           mbm.py:203-211
    temb  = [cos(t f), sin(t f)],  f = exp(-ln(1e4) j / half)  architectures/utils.py:183-198
    w     = exp(-S gamma (1 - t)); B = (w S)/(1 - w); C = w    bridges.py:125-130
    sp    = e^{-g t} (1 - e^{g (t-1)}) / (1 - e^{-g})          bridges.py:218-231 (absorbing only)
"""
import ctypes
import math
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class StepTable:
    n_steps: int
    dt: float            # python float holding the fp32 value
    t: torch.Tensor      # [n_steps] f32 (CPU, contiguous)
    temb: torch.Tensor   # [n_steps, T]
    bc: torch.Tensor     # [n_steps]
    cc: torch.Tensor     # [n_steps]
    sp: Optional[torch.Tensor]  # [n_steps] or None


def sinusoidal_time_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """``t`` [B] f32 -> [B, dim]  (SinusoidalPositionalEncoding.forward, utils.py:183-198)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period)
                      * torch.arange(start=0, end=half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def telegraph_coefficients(t: torch.Tensor, vocab_size: int, gamma: float):
    """B(t), C(t) of TelegraphBridge.rate for ``t`` [B] f32  (bridges.py:125-130)."""
    S = vocab_size
    wt = torch.exp(-S * gamma * (1.0 - t))
    return (wt * S) / (1.0 - wt), wt


def survival_probability(t: torch.Tensor, gamma_absorb: float) -> torch.Tensor:
    """AbsorbingBridge.survival_probability (bridges.py:218-231); gamma is an fp32 tensor there."""
    g = torch.tensor(gamma_absorb, dtype=torch.float32)
    return torch.exp(-g * t) * (1 - torch.exp(g * (t - 1))) / (1 - torch.exp(-g))


def build_step_table(num_timesteps: int, time_eps: float, vocab_size: int, gamma: float, dim_time_emb: int,
                     gamma_absorb: Optional[float] = None) -> StepTable:
    grid = torch.linspace(0.0, 1.0 - time_eps, num_timesteps)
    dt = (grid[-1] - grid[0]) / (len(grid) - 1)
    # state.time = torch.full((B,1), time.item()): the fp32 grid value survives the round trip
    t = grid[1:].clone().contiguous()
    bc, cc = telegraph_coefficients(t, vocab_size, gamma)
    sp = survival_probability(t, gamma_absorb).contiguous() if gamma_absorb is not None else None
    return StepTable(n_steps=num_timesteps - 1, dt=float(dt.item()), t=t,
                     temb=sinusoidal_time_embedding(t, dim_time_emb).contiguous(),
                     bc=bc.contiguous(), cc=cc.contiguous(), sp=sp)


class CStepTable(ctypes.Structure):
    """ctypes image of ``MmbStepTable`` (include/mmbridge.h)."""

    _fields_ = [("n_steps", ctypes.c_int32), ("dt", ctypes.c_float),
                ("t", ctypes.POINTER(ctypes.c_float)), ("temb", ctypes.POINTER(ctypes.c_float)),
                ("bc", ctypes.POINTER(ctypes.c_float)), ("cc", ctypes.POINTER(ctypes.c_float)),
                ("sp", ctypes.POINTER(ctypes.c_float))]

    @staticmethod
    def from_table(table: StepTable) -> "CStepTable":
        fp = ctypes.POINTER(ctypes.c_float)
        ptr = lambda a: ctypes.cast(a.data_ptr(), fp) if a is not None else fp()
        c = CStepTable(table.n_steps, table.dt, ptr(table.t), ptr(table.temb), ptr(table.bc), ptr(table.cc),
                       ptr(table.sp))
        c._keepalive = table  # the arrays are host memory owned by `table`
        return c
