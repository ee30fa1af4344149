"""Reference frequencies of the source-state sampler (container only):  python tests/golden/make_golden_source.py

sample_noise("GaussNoise") + physics_to_onehot/argmax tokens and sample_masks(target_multiplicity=...) of the reference
(mp/data/particle_clouds/utils.py:222-307), 20 000 jets x 30 slots: token and multiplicity frequencies, moments of the
continuous features.  The draws themselves (torch / numpy global generators) cannot be reproduced on the GPU; the fixture pins
the DISTRIBUTIONS."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402,F401

from multimodal_particles.data.particle_clouds.utils import physics_to_onehot, sample_masks, sample_noise  # noqa: E402


def main():
    torch.manual_seed(501)
    np.random.seed(502)
    J, N = 20000, 30
    cat_probs = [0.35, 0.25, 0.2, 0.12, 0.08]
    hist = np.clip(np.rint(np.random.normal(14, 6, 5000)), 0, N).astype(int)
    cont, disc = sample_noise("GaussNoise", num_jets=J, max_num_particles=N, scale=1.5, cat_probs=cat_probs)
    mask = sample_masks(target_multiplicity=hist, min_num_particles=0, max_num_particles=N, num_jets=J)
    tokens = torch.argmax(physics_to_onehot(disc[..., :-1], disc[..., -1]), dim=-1)
    out = dict(cat_probs=np.array(cat_probs, np.float32), hist=hist.astype(np.int32), scale=np.float32(1.5),
               token_freq=np.bincount(tokens.flatten().numpy(), minlength=8) / tokens.numel(),
               mult_freq=np.bincount(mask.sum((1, 2)).numpy(), minlength=N + 1) / J,
               cont_mean=cont.mean((0, 1)).numpy(), cont_std=cont.std((0, 1)).numpy())
    np.savez_compressed(os.path.join(HERE, "source.npz"), **out)
    print("token freq", np.round(out["token_freq"], 4), "mean mult", float(mask.sum((1, 2)).float().mean()))


if __name__ == "__main__":
    main()
