"""Time the standalone fused update kernel (75 B / particle-step) for the variant selected by
MMB_UPDATE_VARIANT; prints one JSON line.  Used to pick the default (profiles/r01_update_variants.md)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multimodal_particles_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for jets in (4096, 32768, 131072):
    r = bench.time_update_kernel(torch, _native, dev, bench.peaks(), flush, jets=jets)
    print(json.dumps({"variant": os.environ.get("MMB_UPDATE_VARIANT", "24"), "jets": jets, "GBps": round(r["achieved"], 1),
                      "frac": round(r["frac"], 3), "us": round(r["ms_per_launch"] * 1e3, 1)}))
