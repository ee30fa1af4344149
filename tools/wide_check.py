"""Wide EPiC (H = 128) tcgen05 trunk against the fp32 kernel, and its speed: python tools/wide_check.py [B]"""
import sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
from multimodal_particles_b200.multimodal_bridge_matching import MultiModalBridgeMatching
from multimodal_particles_b200.databatch import jetclass_like_databatch
from multimodal_particles_b200.epic import as_u8

dev = torch.device("cuda:0")


def wide_model(seed=0, L=6, G=10, skip=True, head=True, ctx=0):
    cfg = MultimodalBridgeMatchingConfig()
    e = cfg.encoder
    e.dim_hidden_local, e.num_blocks, e.dim_hidden_glob, e.skip_connection, e.add_discrete_head = 128, L, G, skip, head
    cfg.data.dim_context_continuous = ctx
    torch.manual_seed(seed)
    return cfg, MultiModalBridgeMatching(cfg).to(dev)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 37
    for kw in (dict(), dict(L=2, G=16, skip=False, head=False), dict(L=3, G=7, ctx=5)):
        cfg, model = wide_model(**kw)
        native = model.encoder.native_model(dev)
        b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(5))
        x, k, m = b.source_continuous.to(dev), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
        T = cfg.encoder.dim_emb_time + kw.get("ctx", 0)
        temb = torch.randn(B, T, device=dev)
        v0, l0, h0 = native.forward(x, k, m, temb, want_hidden=True, precision="fp32")
        v1, l1, h1 = native.forward(x, k, m, temb, want_hidden=True, precision="bf16")
        torch.cuda.synchronize()
        rel = lambda a, b_: ((a - b_).abs().max() / b_.abs().max()).item()
        print(kw, "rel err v %.4f logits %.4f hidden %.4f" % (rel(v1, v0), rel(l1, l0), rel(h1, h0)), "finite", bool(torch.isfinite(v1).all()))
    cfg, model = wide_model()
    native = model.encoder.native_model(dev)
    for B in (296, 4096):
        b = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(6))
        x, k, m = b.source_continuous.to(dev), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
        temb = torch.randn(1, cfg.encoder.dim_emb_time, device=dev)
        for prec in ("bf16", "fp32"):
            for _ in range(2):
                native.forward(x, k, m, temb, precision=prec)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(3):
                native.forward(x, k, m, temb, precision=prec)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 3
            flop = B * 128 * (2 * 6 * 2 * 128 * 128 + 2 * 16 * 128 + 2 * 128 * 11)
            print(f"B={B} {prec}: {ms:.3f} ms  {B / ms * 1e3:.0f} jet-evaluations/s  {flop / ms / 1e9:.1f} TFLOP/s (all 128 slots)")
