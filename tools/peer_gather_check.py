"""Per-batch exchange of the generated jets at N GPUs: NCCL all-gather against the copy-engine push over peer memory
(`sharding.PeerGather`).  Checks that both deliver the same bytes, then times the C2 generation loop (4096 jets per rank and
step, gather + histogram all-reduce on a side stream under the next generation) with no exchange, with NCCL and with the push.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/peer_gather_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multimodal_particles_b200 import MultiModalBridgeMatching, sharding  # noqa: E402
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402
from multimodal_particles_b200.databatch import jetclass_like_databatch  # noqa: E402
from multimodal_particles_b200.epic import as_u8  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))

B, N, K, W = 4096, 128, 20, 5
cfg = MultimodalBridgeMatchingConfig()
cfg.bridge.num_timesteps = 100
torch.manual_seed(0)
model = MultiModalBridgeMatching(cfg).to(dev)
native = model.encoder.native_model(dev)
table = model.step_table()
batch = jetclass_like_databatch(B, N, generator=torch.Generator().manual_seed(1234 + rank))
mask = as_u8(batch.source_mask.to(dev))
hist = sharding.ValidationHistograms(dev, vocab_size=cfg.data.vocab_size_features)
packs = [sharding.PackedJets(B, N, 3, dev, extra_int64=hist.size).load(batch.source_continuous.to(dev), as_u8(batch.source_discrete.to(dev)), mask)
         for _ in range(W + K)]
out = {"world": world}

# ---- same bytes from both
nccl = [sharding.make_gather(B, N, 3, world, dev, mode="nccl", extra_int64=hist.size) for _ in range(2)]
peer = [sharding.make_gather(B, N, 3, world, dev, mode="auto", extra_int64=hist.size) for _ in range(2)]
out["peer_kind"] = peer[0].kind
for rep in range(3):
    packs[0].bytes.random_(0, 255)
    nccl[rep & 1].gather(packs[0])
    peer[rep & 1].gather(packs[0])
    torch.cuda.synchronize()
    same = bool(torch.equal(nccl[rep & 1].bytes, peer[rep & 1].bytes))
    t = torch.tensor([1.0 if same else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out[f"identical_rep{rep}"] = bool(t.item() > 0)
for p in packs:
    p.load(batch.source_continuous.to(dev), as_u8(batch.source_discrete.to(dev)), mask)

# ---- the generation loop of bench.py with each exchange
main_s = torch.cuda.current_stream()
side = torch.cuda.Stream(device=dev)


def loop(gathers, in_band=True):
    def step(i):
        native.generate(packs[i].x, packs[i].k, mask, table, seed=1, jet_offset=rank * B, precision="auto")
        if gathers is not None:
            done = torch.cuda.Event()
            done.record(main_s)
            with torch.cuda.stream(side):
                side.wait_event(done)
                counts = hist.accumulate(packs[i].x, packs[i].k, mask, out=packs[i].counts if in_band else None)
                gathers[i & 1].gather(packs[i], counts)
                last[0] = counts
    for i in range(W):
        step(i)
    main_s.wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(main_s)
    for i in range(K):
        step(W + i)
    main_s.wait_stream(side)
    e.record(main_s)
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for p in packs:
        p.load(batch.source_continuous.to(dev), as_u8(batch.source_discrete.to(dev)), mask)
    return float(t.item()) / K


last = [None]
for name, g, ib in (("none", None, True), ("nccl", nccl, True), ("peer_nccl_counts", peer, False), ("peer", peer, True),
                    ("none2", None, True), ("nccl2", nccl, True), ("peer_nccl_counts2", peer, False), ("peer2", peer, True)):
    ms = loop(g, ib)
    if g is not None:   # every exchange must deliver the same summed counts (one multiplicity entry per jet and rank)
        out[f"jets_in_counts_{name}"] = int(last[0][-(N + 1):].sum().item())
    out[f"ms_per_step_{name}"] = round(ms, 4)
    out[f"M_jets_per_s_{name}"] = round(world * B / ms / 1e3, 3)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
