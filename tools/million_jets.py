"""BASELINE config 5: a 1 M-jet generation run sharded over the GPUs of one box, entirely on the device.

    python tools/million_jets.py [--jets 1048576]                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/million_jets.py

Each rank owns a contiguous slice of the jets (sharding.shard_range; its start is the Philox jet offset, so the jets do not
depend on the number of GPUs) and walks it in micro-batches of 4096: source state on the device (mmb_sample_source), the fused
99-step generation (mmb_generate), post-processing + jet observables (mmb_jet_observables) and the validation histograms
(mmb_validation_histograms) accumulated into per-GPU int64 counts.  Collectives, once per micro-batch and overlapped with the
next one on a side stream: ONE NCCL all-gather of the packed micro-batch (1 792 B / jet) and, at the end, one all-reduce of the
histograms and of the jet-observable sums.  Prints one JSON line on rank 0 (device time by CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from multimodal_particles_b200 import sharding  # noqa: E402
from multimodal_particles_b200.observables import jet_observables  # noqa: E402
from multimodal_particles_b200.source import sample_source_state  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jets", type=int, default=1 << 20)
    ap.add_argument("--micro-batch", type=int, default=4096)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, model = bench.build_model(dev)
    native = model.encoder.native_model(dev)
    table = model.step_table()
    lo, hi = sharding.shard_range(args.jets, rank, world)
    MB, N = args.micro_batch, bench.N_PART
    mult_hist = np.clip(np.rint(np.random.default_rng(0).normal(45, 18, 20000)), 1, N).astype(int)     # JetClass-like multiplicities
    hist = sharding.ValidationHistograms(dev, vocab_size=cfg.data.vocab_size_features)
    counts = torch.zeros(hist.size, dtype=torch.int64, device=dev)
    jet_sums = torch.zeros(11, dtype=torch.float64, device=dev)
    stats = {"mean": [1.2, 0.0, 0.0], "std": [0.35, 0.2, 0.2]}
    packs = [sharding.PackedJets(MB, N, 3, dev) for _ in range(3)]      # state of a micro-batch: [x | tokens | mask], one allocation
    recv = [sharding.PackedGather(MB, N, 3, world, dev) for _ in range(2)] if world > 1 else None
    side = torch.cuda.Stream(device=dev)
    main_s = torch.cuda.current_stream(dev)

    def run(n_jets_from, n_jets_to):
        pending = None
        for i, start in enumerate(range(n_jets_from, n_jets_to, MB)):
            B = min(MB, n_jets_to - start)
            pk = packs[i % 3]
            if B == MB:
                x, k, m = sample_source_state(B, N, target_multiplicity=mult_hist, min_num_particles=0, seed=7, jet_offset=start, compact=True,
                                              out=(pk.x, pk.k, pk.mask))
            else:
                x, k, m = sample_source_state(B, N, target_multiplicity=mult_hist, min_num_particles=0, seed=7, jet_offset=start, compact=True)
            native.generate(x, k, m, table, seed=11, jet_offset=start, precision="bf16")
            _, _, jets = jet_observables(x, k, m, stats, want_particles=False)
            counts.add_(hist.accumulate(x, k, m))
            jet_sums.add_(torch.nan_to_num(jets.double()).sum(0))
            if world > 1 and B == MB:   # gather this micro-batch while the next one is generated
                ev = torch.cuda.Event(); ev.record(main_s)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    recv[i & 1].gather(pk)          # one all-gather of the packed micro-batch
                    pk.bytes.record_stream(side)
                pending = side
        if pending is not None:
            main_s.wait_stream(side)

    run(lo, min(hi, lo + 2 * MB))                    # warm-up (2 micro-batches), then reset the accumulators
    counts.zero_(); jet_sums.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(main_s)
    run(lo, hi)
    if world > 1:
        dist.all_reduce(counts); dist.all_reduce(jet_sums)
    e.record(main_s)
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        off = 3 * hist.bins + cfg.data.vocab_size_features
        mult = counts[off:].cpu().numpy()
        print(json.dumps({"workload": f"C5: {args.jets} jets, N=128, 99 solver steps, source + generation + observables + histograms on the device",
                          "n_gpus": world, "micro_batch": MB, "seconds": ms * 1e-3, "value": args.jets / (ms * 1e-3), "unit": "jets/s",
                          "jets_in_histogram": int(mult.sum()), "mean_multiplicity": float((mult * np.arange(len(mult))).sum() / max(mult.sum(), 1)),
                          "mean_jet_pt": float(jet_sums[4].item() / args.jets), "mean_jet_mass": float(jet_sums[5].item() / args.jets),
                          "token_counts": counts[3 * hist.bins:off].cpu().tolist()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
