"""Per-phase latency of one jet in the tcgen05 transformer-stack kernel (MMB_STACK_TRACE=1): first jet of CTA 0."""
import ctypes, os, sys
os.environ["MMB_STACK_TRACE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_particles_b200 import _native
from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
dev = torch.device("cuda:0")
cfg = AbsorbingConfig(); cfg.data.max_num_particles = 128
torch.manual_seed(0)
gen = AbsorbingFlow(cfg).to(dev).generator
head = gen.native_head(dev)
names = {1: "proj_in row + barrier", 2: "proj_in GEMM round trip", 3: "GN1+swish -> A", 4: "conv1 GEMM round trip", 5: "GN2+swish -> A",
         6: "conv2 GEMM round trip", 7: "GN3 -> A", 8: "fused q|k GEMM round trip", 9: "Q, K tiles", 11: "S GEMM round trip (v issued behind it)",
         12: "softmax -> P, V tile", 13: "PV GEMM round trip", 14: "O tile", 15: "proj_out GEMM round trip", 16: "block 1 (all phases)",
         17: "outputs + barrier"}
for B, packed in ((1, False), (1184, False), (1184, True)):
    hid = torch.randn(B, 128, 16, device=dev); m = torch.ones(B, 128, dtype=torch.uint8, device=dev)
    if packed:   # two 40-particle jets per tile
        m = (torch.arange(128, device=dev)[None] < 40).to(torch.uint8).expand(B, 128).contiguous()
        hid = hid * m[..., None]
    tb = gen.time_bias(torch.tensor([0.5])).to(dev)
    for _ in range(2):
        head.forward(hid, m, tb, pack=packed)
    buf = (ctypes.c_longlong * 48)()
    _native.load().mmb_debug_read_stack_trace(buf, 48)
    t = list(buf)
    print(f"--- B={B} packed={packed}: first tile of CTA 0 = {t[17] - t[0]} cycles")
    prev = t[0]
    for i in sorted(names):
        print(f"  {names[i]:36s} {t[i] - prev:7d}")
        prev = t[i]
    sm = {30: "max pass (2 tcgen05.ld)", 31: "max exchange barrier", 32: "exp pass (2 tcgen05.ld, P tile)", 33: "V tile", 34: "fences", 12: "block barrier"}
    print("  inside the softmax phase of block 0:")
    prev = t[11]
    for i in (30, 31, 32, 33, 34, 12):
        print(f"    {sm[i]:36s} {t[i] - prev:7d}")
        prev = t[i]
