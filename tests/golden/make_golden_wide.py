"""Fixture at the CLASS-DEFAULT widths of the reference's EPiCNetwork (architectures/epic.py:99-101: num_blocks = 6,
dim_hidden_local = 128, dim_hidden_global = 10), from the unmodified reference (container only):

    python tests/golden/make_golden_wide.py

config-mbm-test.yaml with the three encoder widths replaced; 12 time steps, three jets of 128 / 45 / 7 particles in 128 slots.
Same recording as ``make_golden.mbm_case`` (trajectory with injected jump uniforms, heads at selected steps).  This is the shape
the 128-wide tcgen05 trunk (csrc/epic_wide_tc.cu) is built for.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the import shim)

from multimodal_particles.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402


def main():
    cfg = MultimodalBridgeMatchingConfig.from_yaml("/root/reference/tests/resources/configs_files/config-mbm-test.yaml")
    e = cfg.encoder
    e.num_blocks, e.dim_hidden_local, e.dim_hidden_glob = 6, 128, 10
    cfg.bridge.num_timesteps = 12
    B, N, S = 3, 128, 8
    g = torch.Generator().manual_seed(41)
    mask = mg.prefix_masks([128, 45, 7], N)
    x0 = torch.randn(B, N, 3, generator=g) * mask
    k0 = torch.randint(0, S, (B, N, 1), generator=g) * mask
    mg.mbm_case("mbm_wide", cfg, x0, k0, mask, seed=141, snap_steps={0, 5, 10})


if __name__ == "__main__":
    main()
