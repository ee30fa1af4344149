"""BASELINE config 5: a generation run of many jets sharded over the GPUs of one box, entirely on the device.

Each rank owns a contiguous slice of the jets (``sharding.shard_range``; its start is the Philox jet offset, so the jets — and
therefore every histogram count — do not depend on the number of GPUs) and walks it in micro-batches: source state on the
device (``mmb_sample_source``; reference ``sample_noise`` / ``sample_masks``, mp/data/particle_clouds/utils.py:222-286), the
fused solver loop (``mmb_generate``; mbm.py:199-216), post-processing + jet observables (``mmb_jet_observables``;
particles.py:85-89,124-156, jets.py:90-107) and the validation histograms (``mmb_validation_histograms``) accumulated into
per-GPU int64 counts.  Exchanges: the packed micro-batch (1 792 B / jet) goes to every rank once per micro-batch,
on a side stream under the next micro-batch's generation — pushed over NVLink peer memory by the copy engines
(``sharding.PeerGather``; one NCCL all-gather where peer memory is unavailable) — and one all-reduce of the histograms and of the
jet-observable sums at the end.  The reference has no multi-GPU code; ``tools/million_jets.py`` and the
multi-GPU arms of ``bench.py`` call this.
"""
import hashlib

import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from .observables import jet_observables
from .source import sample_source_state

STATS = {"mean": [1.2, 0.0, 0.0], "std": [0.35, 0.2, 0.2]}   # de-standardisation to a jet-like scale (pT > 0)


def sharded_generation_run(model, cfg, total_jets, rank, world, device, micro_batch=4096, n_particles=128, precision="auto",
                           warmup_batches=2, gather="auto"):
    """-> dict (rank 0: the C5 record; other ranks: None).  Device time by CUDA events around the whole slice incl. the final
    all-reduce, max over ranks."""
    native = model.encoder.native_model(device)
    table = model.step_table()
    lo, hi = sharding.shard_range(total_jets, rank, world)
    MB, N = micro_batch, n_particles
    mult_hist = np.clip(np.rint(np.random.default_rng(0).normal(45, 18, 20000)), 1, N).astype(int)     # JetClass-like multiplicities
    hist = sharding.ValidationHistograms(device, vocab_size=cfg.data.vocab_size_features)
    counts = torch.zeros(hist.size, dtype=torch.int64, device=device)
    jet_sums = torch.zeros(11, dtype=torch.float64, device=device)
    packs = [sharding.PackedJets(MB, N, 3, device) for _ in range(3)]      # state of a micro-batch: [x | tokens | mask], one allocation
    recv = [sharding.make_gather(MB, N, 3, world, device, mode=gather) for _ in range(2)] if world > 1 else None
    side = torch.cuda.Stream(device=device)
    main_s = torch.cuda.current_stream(device)

    def run(n_from, n_to):
        # launch stream: source state + solver loop of micro-batch i, back to back; side stream: observables, histograms and the
        # exchange of micro-batch i under the generation of i + 1 (their small kernels find room in its tail).  A packed buffer
        # is refilled only after the side stream is done with it.
        free = [None] * len(packs)
        for i, start in enumerate(range(n_from, n_to, MB)):
            B = min(MB, n_to - start)
            pk = packs[i % 3]
            if free[i % 3] is not None:
                main_s.wait_event(free[i % 3])
            out = (pk.x, pk.k, pk.mask) if B == MB else None
            x, k, m = sample_source_state(B, N, target_multiplicity=mult_hist, min_num_particles=0, device=device, seed=7, jet_offset=start,
                                          compact=True, out=out)
            native.generate(x, k, m, table, seed=11, jet_offset=start, precision=precision)
            ev = torch.cuda.Event()
            ev.record(main_s)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                _, _, jets = jet_observables(x, k, m, STATS, want_particles=False)
                counts.add_(hist.accumulate(x, k, m))
                jet_sums.add_(torch.nan_to_num(jets.double()).sum(0))
                if world > 1 and B == MB:
                    recv[i & 1].gather(pk)          # the packed micro-batch to every rank (copy-engine push, or one all-gather)
                for t in (x, k, m):
                    t.record_stream(side)
                free[i % 3] = torch.cuda.Event()
                free[i % 3].record(side)
        main_s.wait_stream(side)

    run(lo, min(hi, lo + warmup_batches * MB))       # warm-up, then reset the accumulators
    counts.zero_()
    jet_sums.zero_()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(main_s)
    run(lo, hi)
    if world > 1:
        dist.all_reduce(counts)
        dist.all_reduce(jet_sums)
    e.record(main_s)
    torch.cuda.synchronize(device)
    t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms = float(t.item())
    off = 3 * hist.bins + cfg.data.vocab_size_features
    host_counts = counts.cpu().numpy()
    mult = host_counts[off:]
    return {"workload": f"C5: {total_jets} jets, N={N}, {table.n_steps} solver steps; source + generation + observables + histograms on "
                        f"the device, per-micro-batch all-gather of the jets + final all-reduce of the histograms",
            "n_gpus": world, "micro_batch": MB, "exchange": recv[0].kind if recv else "none", "seconds": ms * 1e-3, "value": total_jets / (ms * 1e-3), "unit": "jets/s",
            "precision": native.generate_precision(N, precision),
            "jets_in_histogram": int(mult.sum()), "mean_multiplicity": float((mult * np.arange(len(mult))).sum() / max(mult.sum(), 1)),
            "mean_jet_pt": float(jet_sums[4].item() / total_jets), "mean_jet_mass": float(jet_sums[5].item() / total_jets),
            "token_counts": host_counts[3 * hist.bins:off].tolist(),
            # integer counts of jets keyed by their GLOBAL index: identical for every number of GPUs
            "histogram_sha1": hashlib.sha1(host_counts.astype(np.int64).tobytes()).hexdigest()}
