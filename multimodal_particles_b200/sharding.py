"""Multi-GPU layer of the generation path: jets are independent, so ranks
generate disjoint slices with no data-path collective.  The only exchanges happen once per batch,
after generation: an all-gather of the generated jets in the compact layout (fp32 features, uint8
tokens, uint8 masks = 1 792 B / jet at N=128; one collective when the state lives in a ``PackedJets``
allocation) and an all-reduce (SUM, int64) of the validation histograms.  One process per GPU; ``torch.distributed`` is the plumbing (NCCL on GPUs, gloo in the
CPU tests).  The reference has no distributed code at all: this layer is new.
"""
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ``total`` jets for ``rank``; sizes differ by at most one and the
    slices tile [0, total).  ``lo`` is also the rank's Philox ``jet_offset``, which makes the
    generated jets independent of ``world``."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


RANK_STRIDE = 1 << 40   # jets a single rank can draw before its Philox indices would reach the next rank's block


def global_rank() -> int:
    """Rank of this process in the job: torch.distributed when initialised, else torchrun's / Lightning's environment."""
    import os
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    for name in ("RANK", "GLOBAL_RANK", "SLURM_PROCID"):
        if os.environ.get(name, "").isdigit():
            return int(os.environ[name])
    return 0


def next_jet_offset(owner, count: int, attr: str = "_jets_generated") -> int:
    """Default Philox ``jet_offset`` of a call that draws ``count`` jets: ``rank * 2**40 + jets this process drew before``.
    One process per GPU (Lightning predict / DDP) therefore never reuses a (seed, jet index) pair across ranks; a caller
    that shards explicitly passes its own offset (``shard_range``) and gets results independent of the GPU count."""
    done = getattr(owner, attr, 0)
    setattr(owner, attr, done + count)
    return global_rank() * RANK_STRIDE + done


@dataclass
class GatherBuffers:
    """Preallocated receive buffers for the per-batch all-gather (no allocation in the loop)."""

    x: torch.Tensor
    k: torch.Tensor
    mask: torch.Tensor

    def __init__(self, batch: int, n: int, dim_continuous: int, world: int, device):
        self.x = torch.empty(world * batch, n, dim_continuous, device=device, dtype=torch.float32)
        self.k = torch.empty(world * batch, n, device=device, dtype=torch.uint8)
        self.mask = torch.empty(world * batch, n, device=device, dtype=torch.uint8)


def counts_offset(state_bytes: int) -> int:
    return (state_bytes + 7) & ~7


def packed_bytes(batch: int, n: int, dim_continuous: int, extra_int64: int = 0) -> int:
    state = batch * n * dim_continuous * 4 + 2 * batch * n
    return counts_offset(state) + 8 * extra_int64 if extra_int64 else state


class PackedJets:
    """One contiguous allocation for the compact state of a batch — [x fp32 | tokens u8 | mask u8], 14 B per particle — so the
    per-batch exchange is ONE all-gather instead of three.  ``x`` / ``k`` / ``mask`` are views the kernels work on in place."""

    def __init__(self, batch: int, n: int, dim_continuous: int, device, extra_int64: int = 0):
        self.batch, self.n, self.dc = batch, n, dim_continuous
        nx, nk = batch * n * dim_continuous * 4, batch * n
        self.state_bytes, self.extra_int64 = nx + 2 * nk, extra_int64
        self.bytes = torch.zeros(packed_bytes(batch, n, dim_continuous, extra_int64), dtype=torch.uint8, device=device)
        self.x = self.bytes[:nx].view(torch.float32).view(batch, n, dim_continuous)
        self.k = self.bytes[nx:nx + nk].view(batch, n)
        self.mask = self.bytes[nx + nk:nx + 2 * nk].view(batch, n)
        # optional trailing int64 region (8-byte aligned): the batch's validation-histogram counts travel with the jets
        self.counts = self.bytes[counts_offset(self.state_bytes):].view(torch.int64) if extra_int64 else None

    def load(self, x, k_u8, mask_u8):
        self.x.copy_(x), self.k.copy_(k_u8), self.mask.copy_(mask_u8)
        return self


class PackedGather:
    """Receive side of the packed all-gather: ``world`` PackedJets-shaped slices of one buffer."""

    def __init__(self, batch: int, n: int, dim_continuous: int, world: int, device, extra_int64: int = 0):
        self.world, self.batch, self.n, self.dc = world, batch, n, dim_continuous
        self.nx, self.nk = batch * n * dim_continuous * 4, batch * n
        self.bytes = torch.empty(world, packed_bytes(batch, n, dim_continuous, extra_int64), dtype=torch.uint8, device=device)

    def gather(self, packed: PackedJets, counts=None):
        """One all-gather of the packed jets, one all-reduce (SUM) of the histogram counts (in place)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            self.bytes[0].copy_(packed.bytes)
            return counts
        dist.all_gather_into_tensor(self.bytes.view(-1), packed.bytes)
        if counts is not None:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        return counts

    def x(self, rank: int) -> torch.Tensor:
        return self.bytes[rank, :self.nx].view(torch.float32).view(self.batch, self.n, self.dc)

    def k(self, rank: int) -> torch.Tensor:
        return self.bytes[rank, self.nx:self.nx + self.nk].view(self.batch, self.n)

    def mask(self, rank: int) -> torch.Tensor:
        return self.bytes[rank, self.nx + self.nk:self.nx + 2 * self.nk].view(self.batch, self.n)


class PeerGather(PackedGather):
    """The same exchange without a collective kernel: every rank PUSHES its packed jets into its slot of every peer's receive
    buffer with plain device-to-device copies over NVLink / NVSwitch peer memory (``torch.distributed._symmetric_memory``:
    CUDA VMM allocations mapped into every process of the node), which the COPY ENGINES execute — no SM is involved.  Why: the
    generation kernel is persistent and fills every SM (two CTAs x 256 threads x 128 registers = the whole register file), so an
    NCCL all-gather issued on a side stream cannot start before that kernel's tail, and then holds SMs while the next
    generation kernel starts short of CTAs: at 8 GPUs (58.7 MB received per rank and 4096-jet batch) a solver batch took 1.14 ms
    instead of 1.01 ms.  A copy-engine push overlaps for free.  Completion: one signal-pad barrier (a one-CTA kernel) after the
    pushes — when ``gather`` has run on a rank's stream, the slots of ALL ranks are complete in its buffer.  Contract for re-use:
    a receive buffer may be read on the gathering stream until the next gather INTO THE SAME BUFFER is issued on that stream,
    provided callers alternate at least two buffers (the barrier of gather i + 1 orders every rank's reads of buffer i before
    any push of gather i + 2).  Histogram counts that live in the packed allocation's trailing int64 region (``PackedJets.counts``)
    travel with the same pushes and are summed locally after the barrier — no NCCL kernel in the loop at all; counts held
    elsewhere go through one (tiny) NCCL all-reduce."""

    def __init__(self, batch: int, n: int, dim_continuous: int, world: int, device, extra_int64: int = 0):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.batch, self.n, self.dc = world, batch, n, dim_continuous
        self.nx, self.nk = batch * n * dim_continuous * 4, batch * n
        nbytes = packed_bytes(batch, n, dim_continuous, extra_int64)
        self.rank = dist.get_rank()
        self.bytes = symm_mem.empty((world, nbytes), dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.bytes, dist.group.WORLD)
        # my slot in every rank's buffer, nearest peer first so that the eight ranks do not all write to the same GPU at once
        self.slots = []
        for i in range(world):
            peer = (self.rank + i) % world
            buf = self.bytes if peer == self.rank else self.handle.get_buffer(peer, (world, nbytes), torch.uint8)
            self.slots.append(buf[self.rank])
        self.counts_all = self.bytes[:, counts_offset(self.nx + 2 * self.nk):].view(torch.int64) if extra_int64 else None   # [world, extra]

    def gather(self, packed: PackedJets, counts=None):
        for slot in self.slots:
            slot.copy_(packed.bytes, non_blocking=True)      # cudaMemcpyAsync device-to-device: copy engine
        self.handle.barrier()
        if counts is not None:
            in_band = (self.counts_all is not None and packed.counts is not None and counts.data_ptr() == packed.counts.data_ptr()
                       and counts.numel() == packed.counts.numel())
            if in_band:
                torch.sum(self.counts_all, dim=0, out=counts)   # every rank's counts arrived with its jets
            else:
                dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        return counts


def make_gather(batch: int, n: int, dim_continuous: int, world: int, device, mode: str = "auto", extra_int64: int = 0):
    """Receive side of the per-batch exchange: ``"peer"`` (copy-engine push over peer memory), ``"nccl"`` (one all-gather), or
    ``"auto"``: peer memory when the job runs on CUDA devices of one node under NCCL and the symmetric-memory rendezvous works,
    else the collective.  ``.kind`` says which one was built (``bench.py`` reports it)."""
    device = torch.device(device)
    want_peer = mode == "peer" or (mode == "auto" and device.type == "cuda" and world > 1 and dist.is_available() and dist.is_initialized()
                                   and dist.get_backend() == "nccl")
    if want_peer:
        g, reason = None, ""
        try:
            g = PeerGather(batch, n, dim_continuous, world, device, extra_int64)
        except Exception as exc:   # no symmetric memory on this box / build: the collective does the same job
            if mode == "peer":
                raise
            reason = f"{type(exc).__name__}: {exc}"[:200]
        ok = torch.tensor([1.0 if g is not None else 0.0], device=device)   # every rank must take the same branch
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() > 0:
            g.kind = "peer-push (device-to-device copies over NVLink peer memory on the copy engines + signal barrier)"
            return g
        g = PackedGather(batch, n, dim_continuous, world, device, extra_int64)
        g.kind = f"nccl all-gather (peer memory unavailable: {reason or 'on another rank'})"
        return g
    g = PackedGather(batch, n, dim_continuous, world, device, extra_int64)
    g.kind = "nccl all-gather" if device.type == "cuda" else "all-gather"
    return g


class ValidationHistograms:
    """Per-GPU int64 counts that the all-reduce sums: particle-level histograms of the three
    continuous features, the token multiplicities and the particle multiplicity per jet (the
    observables of JetClassHighLevelFeatures that need no clustering, jets.py:90-107)."""

    def __init__(self, device, vocab_size=8, bins=64, lo=-5.0, hi=5.0, max_particles=128):
        self.device, self.vocab_size, self.bins, self.lo, self.hi = device, vocab_size, bins, lo, hi
        self.max_particles = max_particles
        self.size = 3 * bins + vocab_size + (max_particles + 1)

    def accumulate(self, x: torch.Tensor, k_u8: torch.Tensor, mask_u8: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """Counts of one batch; ``out`` (int64[size], e.g. ``PackedJets.counts``) is overwritten and returned when given."""
        if x.is_cuda:  # one kernel (csrc/histograms.cu); the torch ops below are the host-side statement of it
            from . import _native
            counts = out.zero_() if out is not None else torch.zeros(self.size, dtype=torch.int64, device=x.device)
            return _native.validation_histograms(x, k_u8, mask_u8, counts, self.vocab_size, self.bins, self.lo, self.hi,
                                                 self.max_particles)
        live = mask_u8.bool()
        counts = torch.zeros(self.size, dtype=torch.int64, device=x.device)
        scale = float(torch.tensor(self.bins / (self.hi - self.lo), dtype=torch.float32))
        for c in range(x.shape[-1]):
            idx = ((x[..., c][live] - self.lo) * scale).floor().clamp_(0, self.bins - 1).long()
            counts[c * self.bins:(c + 1) * self.bins] += torch.bincount(idx, minlength=self.bins)
        off = 3 * self.bins
        counts[off:off + self.vocab_size] += torch.bincount(k_u8[live].long(), minlength=self.vocab_size)
        off += self.vocab_size
        mult = mask_u8.sum(dim=1, dtype=torch.int64).clamp_(max=self.max_particles)
        counts[off:] += torch.bincount(mult, minlength=self.max_particles + 1)
        if out is not None:
            out.copy_(counts)
            return out
        return counts


def gather_and_reduce(buffers: GatherBuffers, x, k_u8, mask_u8, counts=None):
    """All-gather the generated jets into ``buffers`` and sum ``counts`` over ranks (in place)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        buffers.x.copy_(x), buffers.k.copy_(k_u8), buffers.mask.copy_(mask_u8)
        return counts
    dist.all_gather_into_tensor(buffers.x, x.contiguous())
    dist.all_gather_into_tensor(buffers.k, k_u8.contiguous())
    dist.all_gather_into_tensor(buffers.mask, mask_u8.contiguous())
    if counts is not None:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
