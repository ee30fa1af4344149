"""HBM roofline of the kernels around the generation loop: source state, observables, validation
histograms, bridge sampling, losses, and one Langevin corrector row.  Working sets larger than the 126 MB L2 (131 072 jets x 128
particles), ten launches per CUDA-event pair, median of five pairs.  Algorithmic bytes per particle are the ones DESIGN.md §4.5-4.7
states.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multimodal_particles_b200 import _native  # noqa: E402
from multimodal_particles_b200.observables import jet_observables  # noqa: E402
from multimodal_particles_b200.sharding import ValidationHistograms  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N, S = 128, 8
P = B * N
peak = bench.peaks()["hbm"]


def timed(fn, launches=10, pairs=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(pairs):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(launches):
            fn()
        e.record()
        torch.cuda.synchronize()
        out.append(s.elapsed_time(e) / launches)
    return sorted(out)[len(out) // 2]


rows = {}


def row(name, ms, nbytes, note):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows[name] = {"ms_per_launch": round(ms, 4), "algorithmic_bytes": int(nbytes), "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                  "note": note}


g = torch.Generator(device=dev).manual_seed(1)
import ctypes  # noqa: E402
from multimodal_particles_b200.source import multiplicity_cdf  # noqa: E402
x = torch.empty(B, N, 3, device=dev)
k = torch.empty(B, N, dtype=torch.uint8, device=dev)
mask = torch.empty(B, N, dtype=torch.uint8, device=dev)
mult = torch.randint(1, N + 1, (4096,))
cdf = torch.from_numpy(multiplicity_cdf(mult, N)).to(dev)
probs = (ctypes.c_float * 5)(0.2, 0.2, 0.2, 0.2, 0.2)
lib = _native.load()
row("sample_source", timed(lambda: _native.check(lib.mmb_sample_source(_native._ptr(x), _native._ptr(k), _native._ptr(mask), B, N, 1.0, probs,
                                                                       _native._ptr(cdf), 3, 0, _native._stream()))),
    P * 14, "14 B per particle slot written (x, token, mask); Philox + Box-Muller in-kernel")
live = float(mask.float().mean())
row("jet_observables", timed(lambda: jet_observables(x, k, mask)), P * 28 + B * 44, "28 B per particle + 44 B per jet")
vh = ValidationHistograms(dev, vocab_size=S, max_particles=N)
row("validation_histograms", timed(lambda: vh.accumulate(x, k, mask)), P * 14, "14 B per particle read; int64 counts through shared-memory atomics")

x0, x1 = torch.randn(B, N, 3, device=dev, generator=g), torch.randn(B, N, 3, device=dev, generator=g)
k0 = torch.randint(0, S, (B, N), device=dev, dtype=torch.uint8, generator=g)
k1 = torch.randint(0, S, (B, N), device=dev, dtype=torch.uint8, generator=g)
t = torch.rand(B, device=dev, generator=g)
row("sample_bridges", timed(lambda: _native.sample_bridges(x0, x1, k0, k1, t, 1e-4, 0.075, S, seed=5)), P * 39,
    "39 B per particle; includes the allocation of the two outputs by the wrapper")
row("absorbing_sample", timed(lambda: _native.absorbing_sample(t, mask, seed=5)), P * 2, "2 B per particle")
v, lg = torch.randn(B, N, 3, device=dev, generator=g), torch.randn(B, N, S, device=dev, generator=g)
row("bridge_losses", timed(lambda: _native.bridge_losses(v, lg, x0, x1, k1, mask)), P * 70, "70 B per particle, deterministic two-pass sum")
del x0, x1, k0, k1

# one Langevin corrector row with in-kernel noise: norms pass (44 B per live particle) + the update pair (133 B)
Bc = B // 2
dims = torch.randint(1, N + 1, (Bc,), generator=torch.Generator().manual_seed(2)).to(dev, torch.int32)
m = (torch.arange(N, device=dev)[None] < dims[:, None]).float().unsqueeze(-1)
xs, oh = torch.randn(Bc, N, 3, device=dev, generator=g) * m, torch.randn(Bc, N, S, device=dev, generator=g) * m
vc, lc = v[:Bc] * m, lg[:Bc] * m
n_live = int(dims.sum())
row("corrector_row", timed(lambda: _native.trans_corrector_update(xs, oh, dims, vc, lc, None, None, None, 0.98, True, 1.5, 0.1, 0.001, seed=1)),
    n_live * 177, "norm pass 44 B + update pair 133 B per LIVE particle (mean multiplicity %.1f); three kernels + the 1-block coefficient kernel" % (n_live / Bc))
print(json.dumps({"tool": "aux_bench", "jets": B, "particles_per_jet": N, "source_live_fraction": round(live, 3), "hbm_peak_gbs": peak,
                  "l2": "working sets 235 MB - 1.2 GB > 126 MB L2; 10 launches per event pair, median of 5", "kernels": rows}))
