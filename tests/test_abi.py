"""CPU: the C-ABI library loads without a GPU and exports exactly what include/mmbridge.h declares."""
import ctypes
import os
import re

import pytest

from multimodal_particles_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mmbridge.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(mmb_[a-z0-9_]+)\s*\(", text))
    names.discard("mmb_epic_layout")  # static inline helper
    return names


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = declared_symbols()
    assert len(names) >= 9
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mmbridge.h but not exported"
    assert names == set(_native.SIGNATURES), "binding table and header disagree"
    assert lib.mmb_abi_version() == 3


def test_layout_size_matches_python_packing():
    import torch
    from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import (
        MultimodalBridgeMatchingConfig)
    from multimodal_particles_b200.multimodal_bridge_matching import MultiModalBridgeMatching
    lib = _native.load()
    for tweak in (dict(), dict(dim_hidden_glob=19, num_blocks=3, add_discrete_head=False, dim_emb_time=14)):
        cfg = MultimodalBridgeMatchingConfig()
        for key, val in tweak.items():
            setattr(cfg.encoder, key, val)
        model = MultiModalBridgeMatching(cfg)
        enc = model.encoder
        head = enc.fc_layer if enc.add_discrete_head else None
        dims = enc.epic.epic_dims(head[0].out_features if head is not None else 0)
        assert lib.mmb_epic_packed_floats(ctypes.byref(dims)) == enc.epic.pack_weights(head).numel()


def test_argument_errors_are_codes_not_crashes():
    lib = _native.load()
    rc = lib.mmb_bridge_update(None, None, None, None, None, None, None, None, 0.0, 0.0, 0.0, 0.0, -1, 4, 3, 8, 0, None)
    assert rc == -1 and b"negative" in lib.mmb_last_error()
    rc = lib.mmb_bridge_update(None, None, None, None, None, None, None, None, 0.0, 0.0, 0.0, 0.0, 1, 4, 3, 64, 0, None)
    assert rc == -1
    rc = lib.mmb_epic_forward(None, None, None, None, None, 0, 1, 1, None, None, None, 0, None)
    assert rc == -1


def test_no_cpu_fallback():
    """Host tensors are rejected: the product never routes through a CPU implementation."""
    import torch
    x = torch.zeros(1, 4, 3)
    with pytest.raises(_native.MmbError):
        _native.bridge_update(x, torch.zeros(1, 4, dtype=torch.uint8), torch.ones(1, 4, dtype=torch.uint8),
                              x, torch.zeros(1, 4, 8), torch.zeros(1, 4), 0.01, 1.0, 0.5)
