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
