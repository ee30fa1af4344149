"""Golden fixture of the forward half of a training / validation step, produced by the reference (container only):

    python tests/golden/make_golden_train.py

MultiModalBridgeMatching.sample_bridges, loss_continuous, loss_discrete (mp/models/generative/multimodal_bridge_matching.py:148-197)
and AbsorbingBridge.sample (bridges.py:233-249) with the draws injected: torch.rand -> t, torch.randn_like -> z,
torch.rand_like -> u, and torch.distributions.Categorical.sample replaced by the inverse CDF on injected uniforms
(the same distribution; its own generator stream cannot be reproduced on the GPU)."""
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402,F401

from multimodal_particles.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402
from multimodal_particles.config_classes.absorbing_flows_config import AbsorbingConfig  # noqa: E402
from multimodal_particles.models.generative import bridges  # noqa: E402
from multimodal_particles.models.generative.multimodal_bridge_matching import MultiModalBridgeMatching, MultiHeadOutput  # noqa: E402
from types import SimpleNamespace  # noqa: E402


def main():
    cfg = MultimodalBridgeMatchingConfig()
    cfg.data.max_num_particles = 20
    cfg.bridge.sigma = 0.05
    torch.manual_seed(601)
    model = MultiModalBridgeMatching(cfg)
    g = torch.Generator().manual_seed(602)
    B, N, S = 9, 20, cfg.data.vocab_size_features
    mask = (torch.arange(N)[None] < torch.tensor([1, 3, 20, 8, 2, 5, 17, 11, 20])[:, None]).long().unsqueeze(-1)
    batch = SimpleNamespace(source_continuous=torch.randn(B, N, 3, generator=g), source_discrete=torch.randint(0, S, (B, N, 1), generator=g),
                            source_mask=mask, target_continuous=torch.randn(B, N, 3, generator=g) * mask,
                            target_discrete=torch.randint(0, S, (B, N, 1), generator=g) * mask, target_mask=mask)
    t = torch.rand(B, generator=g)
    t[0], t[1] = 0.003, 0.997
    z = torch.randn(B, N, 3, generator=g)
    u = torch.rand(B, N, generator=g)

    class InjectedCategorical:
        def __init__(self, probs):
            self.probs = probs / probs.sum(-1, keepdim=True)

        def sample(self):
            c = torch.cumsum(self.probs, -1)
            return (u[..., None] >= c).sum(-1).clamp(max=self.probs.shape[-1] - 1)

    with mock.patch.object(torch, "rand", lambda *a, **k: t.clone()), mock.patch.object(torch, "randn_like", lambda x: z.clone()), \
            mock.patch.object(bridges, "Categorical", InjectedCategorical):
        state = model.sample_bridges(batch)
    v, logits = torch.randn(B, N, 3, generator=g), torch.randn(B, N, S, generator=g) * 2
    heads = MultiHeadOutput(v, logits, mask)
    l0, l1 = model.loss_continuous(heads, state, batch), model.loss_discrete(heads, state, batch)
    # absorbing bridge sample
    acfg = AbsorbingConfig()
    ab = bridges.AbsorbingBridge(acfg)
    ua = torch.rand(B, N, 1, generator=g)
    with mock.patch.object(torch, "rand_like", lambda x: ua.clone()):
        mask_t = ab.sample(t.view(B, 1, 1), mask)
    out = dict(sigma=np.float32(cfg.bridge.sigma), gamma=np.float32(cfg.bridge.gamma), x0=batch.source_continuous.numpy(),
               x1=batch.target_continuous.numpy(), k0=batch.source_discrete.numpy().astype(np.uint8),
               k1=batch.target_discrete.numpy().astype(np.uint8), mask=mask.numpy().astype(np.uint8), t=t.numpy(), z=z.numpy(), u=u.numpy(),
               xt=state.continuous.numpy(), kt=state.discrete.numpy().astype(np.uint8), time=state.time.numpy(), v=v.numpy(),
               logits=logits.numpy(), loss_continuous=np.float32(l0.item()), loss_discrete=np.float32(l1.item()),
               gamma_absorb=np.float32(acfg.bridge.gamma_absorb), u_absorb=ua.numpy(), sp=ab.survival_probability(t).numpy(),
               mask_t=mask_t.numpy().astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, "train_forward.npz"), **out)
    print("losses", l0.item(), l1.item(), "tokens moved", int((state.discrete != batch.source_discrete).sum()), "born", int(mask_t.sum() - mask.sum()))


if __name__ == "__main__":
    main()
