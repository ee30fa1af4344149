"""The transformer-stack kernel alone: one row per slot vs padded slots once + several jets per tile (pack), on inputs whose
padded slots are identical (packable: what the trunk produces) and on random inputs (not packable), JetClass-like and uniform
multiplicities.  72.0 MFLOP per jet (absorbing head) -> TFLOP/s against the measured bf16 peak."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.epic import as_u8
dev = torch.device("cuda:0")
cfg = AbsorbingConfig(); cfg.data.max_num_particles = 128
torch.manual_seed(0)
gen = AbsorbingFlow(cfg).to(dev).generator
head = gen.native_head(dev)
B, N = 4096, 128
pk = bench.peaks()
tb = gen.time_bias(torch.tensor([0.5])).to(dev)
g = torch.Generator().manual_seed(3)
masks = {"jetclass-like (mean 45)": as_u8(jetclass_like_databatch(B, N, generator=g).source_mask),
         "uniform 1..128": (torch.arange(N)[None] < torch.randint(1, N + 1, (B, 1), generator=g)).to(torch.uint8),
         "all 128 live": torch.ones(B, N, dtype=torch.uint8)}
out = {}
for mname, m in masks.items():
    m = m.to(dev)
    for dname, hid in (("identical padded slots", torch.randn(B, N, 16, device=dev) * m[..., None]), ("random padded slots", torch.randn(B, N, 16, device=dev))):
        for pack in (False, True):
            if os.environ.get("STACK_BENCH_ONLY") and os.environ["STACK_BENCH_ONLY"] not in f"{mname}|{dname}|pack={pack}":
                continue
            for _ in range(2):
                head.forward(hid, m, tb, pack=pack)
            ts = []
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); head.forward(hid, m, tb, pack=pack); e.record(); torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[2]
            tf = 72.0e6 * B / (ms * 1e-3) / 1e12
            print(f"{mname:26s} {dname:24s} pack={pack!s:5s}: {ms:.3f} ms  {tf:6.1f} TFLOP/s  {tf / pk['bf16']:.3f} of peak")
            out[f"{mname}|{dname}|pack={pack}"] = {"ms": ms, "tflops": tf, "frac": tf / pk["bf16"]}
print(json.dumps(out))
