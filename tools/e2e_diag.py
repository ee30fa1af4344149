import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench
from multimodal_particles_b200 import HybridState
dev = torch.device("cuda:0")
cfg, model = bench.build_model(dev)
batch = bench.source_batch(4096, 1234)
pin = lambda t: t.clone().pin_memory()
states = [HybridState(None, pin(batch.source_continuous), pin(batch.source_discrete), pin(batch.source_mask)) for _ in range(14)]
model.precision = "bf16"
for i in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = model.simulate_dynamics(states[i], batch, jet_offset=0)
    torch.cuda.synchronize(); print(f"{i}: {(time.perf_counter()-t0)*1e3:.2f} ms")
