"""Phase trace of the wide EPiC trunk (first pair of CTA 0): MMB_WIDE_TRACE=1 python tools/wide_trace.py"""
import ctypes, os, sys
os.environ["MMB_WIDE_TRACE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from wide_check import wide_model, dev
from multimodal_particles_b200 import _native
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.epic import as_u8
cfg, model = wide_model()
native = model.encoder.native_model(dev)
names = ["load pair", "local_0 issue + tv0", "epi0(A)", "publish+gemm+epi0(B)", "projection globals", "layer 0 (all)", "layer 1: globals", "epi1(A)",
         "publish+gemm l2(A)", "epi1(B)", "publish+gemm l2(B)", "epi2(A)", "publish+gemm+epi2(B)+gemm", "layers 2..", "output"]
for B in (2, 2368):
    b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(6))
    x, k, m = b.source_continuous.to(dev), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
    temb = torch.randn(1, cfg.encoder.dim_emb_time, device=dev)
    for _ in range(2):
        native.forward(x, k, m, temb, precision="bf16")
    buf = (ctypes.c_longlong * 32)()
    _native.load().mmb_debug_read_wide_trace(buf, 32)
    t = list(buf)
    print(f"--- B={B}: first pair of CTA 0 = {t[15] - t[0]} cycles")
    for i, n in enumerate(names):
        print(f"  {n:32s} {t[i + 1] - t[i]:7d}")
    print("  inside the globals of layer 1: frag loads issued %d | pool + barrier %d | fc_global1 %d | write + barrier %d | fc_global2 + write + barrier %d | fc_local1 part %d | write + barrier %d"
          % (t[16] - t[6], t[17] - t[16], t[18] - t[17], t[19] - t[18], t[20] - t[19], t[21] - t[20], t[7] - t[21]))
