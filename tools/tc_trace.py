"""Per-phase latency of one jet-step of the tcgen05 generation kernel (MMB_TC_TRACE=1)."""
import ctypes, os, sys
os.environ["MMB_TC_TRACE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from multimodal_particles_b200 import _native
from multimodal_particles_b200.epic import as_u8
dev = torch.device("cuda:0")
cfg, model = bench.build_model(dev)
native = model.encoder.native_model(dev)
table = model.step_table()
names = {1: "time vectors + A0 row + barrier", 2: "local_0 MMA round trip", 3: "local_0 epilogue + barrier",
         4: "L0 pool+fc_local1 MMA, global MLP, barrier", 5: "L0 fc_local1 epilogue + barrier", 6: "L0 fc_local2 MMA round trip",
         7: "L0 fc_local2 epilogue + barrier", 8: "L1 pool+fc_local1 MMA, global MLP, barrier", 9: "L1 fc_local1 epilogue + barrier",
         10: "L1 fc_local2 MMA round trip", 11: "L1 fc_local2 epilogue + barrier", 12: "out+head0 MMA round trip",
         14: "selu + barrier + head2 MMA round trip", 15: "logits load", 16: "update (Euler, philox, jump)"}
for B in (1, 1184, 2368):
    b = bench.source_batch(B, 1)
    x, k, m = b.source_continuous.to(dev).contiguous(), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
    for _ in range(2):
        native.generate(x.clone(), k.clone(), m, table, seed=1, precision="bf16")
    buf = (ctypes.c_longlong * 32)()
    _native.load().mmb_debug_read_trace(buf, 32)
    t = list(buf)
    print(f"--- B={B}: one solver step of jet 0 = {t[16] - t[0]} cycles")
    print(f"  special lane: local_0 issue {t[21] - t[20]}, commit->mbarrier {t[22] - t[21]}; L0 fc_local2 issue {t[24] - t[23]}, commit->mbarrier {t[25] - t[24]}")
    prev = t[0]
    for i in sorted(names):
        print(f"  {names[i]:48s} {t[i] - prev:6d}")
        prev = t[i]
