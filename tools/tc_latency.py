"""Chain latency of the tcgen05 generation kernel: time for 99 steps at tiny batches (no contention)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from multimodal_particles_b200.epic import as_u8
dev = torch.device("cuda:0")
cfg, model = bench.build_model(dev)
native = model.encoder.native_model(dev)
table = model.step_table()
for B in (1, 4, 148 * 4, 148 * 8, 148 * 8 + 4, 4096):
    b = bench.source_batch(B, 1)
    x, k, m = b.source_continuous.to(dev).contiguous(), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
    for _ in range(3):
        native.generate(x.clone(), k.clone(), m, table, seed=1, precision="bf16")
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xs = [x.clone() for _ in range(5)]; ks = [k.clone() for _ in range(5)]
    s.record()
    for i in range(5):
        native.generate(xs[i], ks[i], m, table, seed=1, precision="bf16")
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"B={B:5d}: {ms:.3f} ms per generation, {ms * 1e3 / 99:.2f} us per solver step, {B / ms * 1e3:.0f} jets/s")
