"""CPU, world_size 2 over gloo: the host logic of the multi-GPU layer — slice
arithmetic, all-gather layout of the generated jets, all-reduce of the validation histograms, and
sharding invariance of the generation itself (checked with the CPU oracle standing in for the
kernels: the Philox key is the GLOBAL jet index, so two ranks must reproduce the single-rank jets)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from multimodal_particles_b200 import sharding
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.multimodal_bridge_matching import MultiModalBridgeMatching


def test_shard_range_tiles_the_jets():
    for total, world in [(10, 3), (1_000_000, 8), (7, 8), (4096, 2), (0, 4)]:
        edges = [sharding.shard_range(total, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = MultimodalBridgeMatchingConfig()
        cfg.bridge.num_timesteps = 6
        torch.manual_seed(0)
        model = MultiModalBridgeMatching(cfg)
        dims, packed = ol.packed_model(model)
        full = jetclass_like_databatch(total, 32, generator=torch.Generator().manual_seed(9))
        lo, hi = sharding.shard_range(total, rank, world)
        x, k = ol.generate(dims, packed, full.source_continuous[lo:hi].numpy(), full.source_discrete[lo:hi, :, 0].numpy(),
                           full.source_mask[lo:hi, :, 0].numpy(), model.step_table(), seed=3, jet_offset=lo, nthreads=1)
        x, k = torch.from_numpy(x), torch.from_numpy(k)
        m = full.source_mask[lo:hi, :, 0].to(torch.uint8)
        hist = sharding.ValidationHistograms("cpu", vocab_size=8, max_particles=32)
        counts = hist.accumulate(x, k, m)
        buf = sharding.GatherBuffers(hi - lo, 32, 3, world, "cpu")
        counts = sharding.gather_and_reduce(buf, x, k, m, counts)
        # the same exchange as ONE collective on the packed layout [x | tokens | mask]
        pk = sharding.PackedJets(hi - lo, 32, 3, "cpu").load(x, k, m)
        pg = sharding.PackedGather(hi - lo, 32, 3, world, "cpu")
        pg.gather(pk)
        px = torch.cat([pg.x(r) for r in range(world)])
        pk_ = torch.cat([pg.k(r) for r in range(world)])
        pm = torch.cat([pg.mask(r) for r in range(world)])
        assert torch.equal(px, buf.x) and torch.equal(pk_, buf.k) and torch.equal(pm, buf.mask)
        # off CUDA / NCCL the factory hands out the collective, never the peer-memory push
        mg = sharding.make_gather(hi - lo, 32, 3, world, "cpu")
        assert type(mg) is sharding.PackedGather and mg.kind == "all-gather"
        mg.gather(pk)
        assert torch.equal(mg.bytes, pg.bytes)
        if rank == 0:
            np.savez(out_path, x=buf.x.numpy(), k=buf.k.numpy(), mask=buf.mask.numpy(), counts=counts.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_generation_equals_single_rank(tmp_path):
    total, world = 8, 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    z = np.load(out)
    cfg = MultimodalBridgeMatchingConfig()
    cfg.bridge.num_timesteps = 6
    torch.manual_seed(0)
    model = MultiModalBridgeMatching(cfg)
    dims, packed = ol.packed_model(model)
    full = jetclass_like_databatch(total, 32, generator=torch.Generator().manual_seed(9))
    x, k = ol.generate(dims, packed, full.source_continuous.numpy(), full.source_discrete[..., 0].numpy(),
                       full.source_mask[..., 0].numpy(), model.step_table(), seed=3, jet_offset=0)
    assert np.array_equal(z["x"], x) and np.array_equal(z["k"], k)
    assert np.array_equal(z["mask"], full.source_mask[..., 0].numpy())
    hist = sharding.ValidationHistograms("cpu", vocab_size=8, max_particles=32)
    want = hist.accumulate(torch.from_numpy(x), torch.from_numpy(k), full.source_mask[..., 0].to(torch.uint8))
    assert np.array_equal(z["counts"], want.numpy())
    assert int(z["counts"][-33:].sum()) == total  # one multiplicity entry per jet


def _peer_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, N = 37, 128
        peer = [sharding.make_gather(B, N, 3, world, dev, mode="peer", extra_int64=5) for _ in range(2)]
        nccl = sharding.make_gather(B, N, 3, world, dev, mode="nccl", extra_int64=5)
        assert isinstance(peer[0], sharding.PeerGather) and type(nccl) is sharding.PackedGather
        ok = True
        for rep in range(4):   # two receive buffers in turn, fresh bytes every time
            pk = sharding.PackedJets(B, N, 3, dev, extra_int64=5)
            pk.bytes.copy_(torch.randint(0, 256, (pk.bytes.numel(),), dtype=torch.uint8, generator=torch.Generator().manual_seed(100 * rep + rank)))
            pk.counts.fill_(rank + 1 + rep)
            nccl.gather(pk)
            # counts inside the packed allocation travel with the pushes (even reps); counts held elsewhere take the all-reduce
            counts = pk.counts if rep % 2 == 0 else torch.full((5,), rank + 1 + rep, dtype=torch.int64, device=dev)
            peer[rep & 1].gather(pk, counts)
            torch.cuda.synchronize()
            ok = ok and torch.equal(peer[rep & 1].bytes, nccl.bytes) and int(counts[0]) == sum(r + 1 + rep for r in range(world))
            # views of the receive buffer (compared as bytes: random bytes make NaN floats)
            ok = ok and torch.equal(peer[rep & 1].x(rank).view(torch.uint8), pk.x.view(torch.uint8)) and torch.equal(peer[rep & 1].mask(rank), pk.mask)
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            np.savez(out_path, ok=t.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_push_gather_equals_nccl_all_gather(tmp_path):
    """sharding.PeerGather (copy-engine push over NVLink peer memory) delivers the bytes of one NCCL all-gather; needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "peer.npz")
    mp.spawn(_peer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert float(np.load(out)["ok"][0]) == 1.0
