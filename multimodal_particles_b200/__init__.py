"""multimodal_particles_b200 — the B200-native generation hot path of Multimodal-Bridges.

Host-side mirror of the reference's model/sampler API over libmmbridge.so (csrc/).  Only what the
generation path needs lives here: configs, state containers, the EPiC parameter
tree, the bridges' solver steps and the two generators.
"""
from .states import AbsorbingBridgeState, HybridState, MultiHeadOutput, OutputHeads  # noqa: F401
from .multimodal_bridge_matching import MultiModalBridgeMatching  # noqa: F401
