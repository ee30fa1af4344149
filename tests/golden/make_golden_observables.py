"""Golden fixture of the post-processing / jet observables, produced by the reference (container only):

    python tests/golden/make_golden_observables.py

``ParticleClouds(dataset=HybridState)`` -> ``postprocess(stats=...)`` -> ``JetClassHighLevelFeatures`` (kinematics, multiplicity,
jet charges; mp/data/particle_clouds/particles.py:33-38,85-89,124-156, jets.py:85-107,138-141).  The fastjet substructure call
at the end of ``JetClassHighLevelFeatures.__init__`` is disabled (fastjet is not installed here); nothing else is touched.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402,F401  (installs the shim)

from multimodal_particles.data.particle_clouds.particles import ParticleClouds  # noqa: E402
from multimodal_particles.data.particle_clouds.jets import JetClassHighLevelFeatures  # noqa: E402
from multimodal_particles.models.generative.multimodal_bridge_matching import HybridState  # noqa: E402


def main():
    g = torch.Generator().manual_seed(401)
    B, N = 12, 40
    mult = torch.tensor([0, 1, 2, 5, 9, 13, 20, 27, 33, 39, 40, 40])
    mask = (torch.arange(N)[None] < mult[:, None]).long().unsqueeze(-1)
    x = torch.randn(B, N, 3, generator=g) * mask
    k = torch.randint(0, 8, (B, N, 1), generator=g) * mask
    stats = {"mean": [3.1, -0.02, 0.01], "std": [6.5, 0.21, 0.23]}
    pc = ParticleClouds(dataset=HybridState(None, x.clone(), k.clone(), mask.clone()))
    pc.postprocess(input_continuous="standardize", input_discrete="tokens", stats=stats)
    JetClassHighLevelFeatures.substructure = lambda self: None
    jets = JetClassHighLevelFeatures(pc)
    cols = ["px", "py", "pz", "e", "pt", "m", "eta", "phi"]
    out = dict(x=x.numpy(), k=k.numpy().astype(np.uint8), mask=mask.numpy().astype(np.uint8), mean=np.array(stats["mean"], np.float32),
               std=np.array(stats["std"], np.float32), continuous=pc.continuous.numpy(), flavor=pc.flavor.numpy().astype(np.int8),
               charge=pc.charge.numpy().astype(np.int8), discrete=pc.discrete.numpy().astype(np.int8),
               multiplicity=jets.multiplicity.numpy().astype(np.int32), Q_total=jets.Q_total.numpy(), Q_jet=jets.Q_jet.numpy())
    for c in cols:
        out[f"jet_{c}"] = getattr(jets, c).numpy()
    path = os.path.join(HERE, "observables.npz")
    np.savez_compressed(path, **out)
    print(f"observables: {os.path.getsize(path) / 1024:.0f} KiB; jet pt {jets.pt[:4].tolist()} m {jets.m[:4].tolist()}")


if __name__ == "__main__":
    main()
